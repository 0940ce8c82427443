"""Free-function surface of the reference's `group/su3/pytorch/utils.py`, so that
`from l2hmc.group.su3.pytorch.utils import projectSU, su3_to_vec, ...` keeps working after the import swap.

Only the functions of SURVEY section 8 a18 -- projectSU / projectU / rsqrtPHM3 / rsqrtPHM3f / eigs3x3,
su3_to_vec / vec_to_su3, randTAH3, norm2, eyeOf, checkSU / checkU, projectTAH -- live here: implemented in `group.py`
on top of the libl2b kernels and re-exported, plus `eigs3x3` as the free function the reference exposes.  The
reference's other helpers in that file (`cmax`, `unit`, `charpoly3x3`, `expm`, `cubic_zeros`, `log3x3`, ...) are
not on the hot path and are not mirrored.
"""
from __future__ import annotations

import torch

from .group import (checkSU, checkU, eyeOf, norm2, projectSU, projectTAH, projectU, randTAH3, rsqrtPHM3,  # noqa: F401
                    rsqrtPHM3f, su3_to_vec, vec_to_su3)

Tensor = torch.Tensor


def eigs3x3(tr: Tensor, p2: Tensor, det: Tensor) -> tuple[Tensor, Tensor, Tensor]:
    """eigenvalues of a Hermitian 3x3 from tr X, tr X^2, det X by the trigonometric closed form, with the
    reference's clamps (utils.py:227-283); the same formula runs per link inside the projectSU kernels"""
    tr3, tr32 = tr / 3.0, (tr / 3.0) ** 2
    q = (0.5 * (p2 / 3.0 - tr32)).abs()
    r = 0.25 * tr3 * (5.0 * tr32 - p2) - 0.5 * det
    sq = q.sqrt()
    isq3 = (1.0 / (q * sq)).clamp(-3e38, 3e38)
    rsq3 = (r * isq3).clamp(-1.0, 1.0).clamp(-1.0 + 1e-12, 1.0 - 1e-12)
    t = torch.acos(rsq3) / 3.0
    sqc, sqs = sq * t.cos(), (3.0 ** 0.5) * sq * t.sin()
    ll = tr3 + sqc
    return tr3 - 2.0 * sqc, ll + sqs, ll - sqs

"""Free-function surface of the reference's `group/su3/pytorch/utils.py`, so that
`from l2hmc.group.su3.pytorch.utils import projectSU, su3_to_vec, ...` keeps working after the import swap.

The functions the integrator uses (projectSU / projectU / rsqrtPHM3 / rsqrtPHM3f / eigs3x3,
su3_to_vec / vec_to_su3, randTAH3, norm2, eyeOf, checkSU / checkU, projectTAH) are implemented in
`group.py` on top of the libl2b kernels and re-exported here; the small algebra helpers below are plain torch.
Not mirrored: the reference's unused experiments (`cubic_zeros`, `su3_to_eigs`, `log3x3`, `acos_safe*`,
`su3fabc`, `SU3Gradient`) -- nothing in the reference calls them.
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from .group import (checkSU, checkU, eyeOf, norm2, projectSU, projectTAH, projectU, randTAH3, rsqrtPHM3,  # noqa: F401
                    rsqrtPHM3f, su3_to_vec, vec_to_su3)

Tensor = torch.Tensor
eyeOf1 = eyeOf                      # utils.py:125-131: same identity, older spelling


def cmax(x: Tensor, y: Tensor) -> Tensor:
    """element-wise, the argument of larger magnitude (utils.py:50-52)"""
    return torch.where(x.abs() > y.abs(), x, y)


def unit(shape: Sequence[int], dtype: Optional[torch.dtype] = torch.complex128) -> Tensor:
    """an identity broadcastable against a batch of `shape[-2:]` matrices (utils.py:55-62)"""
    eye = torch.eye(int(shape[-1]), dtype=dtype)
    return eye.reshape(*([1] * (len(shape) - 2)), *eye.shape)


def eye_like(x: Tensor) -> Tensor:
    """identity with x's (2-d) shape, dtype and device (utils.py:144-145)"""
    return torch.eye(*x.shape, dtype=x.dtype, device=x.device)


def charpoly3x3(a: Tensor) -> tuple[Tensor, Tensor, Tensor]:
    """det(l - A) = l^3 + c3 l^2 + c2 l + c1 for a batch [n, 3, 3]; returns (c1, c2, c3) = (-det A,
    sum of principal 2x2 minors, -tr A)   (utils.py:65-82)"""
    tr = torch.diagonal(a, dim1=-2, dim2=-1).sum(-1)
    tr2 = torch.diagonal(a @ a, dim1=-2, dim2=-1).sum(-1)
    return -torch.linalg.det(a), 0.5 * (tr * tr - tr2), -tr


def expm(m: Tensor, order: int = 12) -> Tensor:
    """Taylor polynomial of exp(m) of degree `order` in Horner form (utils.py:148-154) -- a truncated series, not
    `torch.matrix_exp`; the integrator itself uses SU3.exp (the Cayley-Hamilton kernel `l2b_su3_exp`)"""
    eye = torch.eye(m.shape[-1], dtype=m.dtype, device=m.device)
    x = eye + m / order
    for i in range(order - 1, 0, -1):
        x = eye + (m @ x) / i
    return x


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

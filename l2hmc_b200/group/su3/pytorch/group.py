"""`SU3` group object with the reference's method surface
(`group/su3/pytorch/group.py:36-227`, `group/su3/pytorch/utils.py`), every
arithmetic method backed by an sm_100a kernel of libl2b (include/l2b.h).

Tensors are complex128 CUDA tensors `[..., 3, 3]`; field-shaped calls use the
reference layout `[nb, 4, T, X, Y, Z, 3, 3]`.
"""
from __future__ import annotations

import itertools
from typing import Optional, Sequence

import torch

from ....group.group import Group
from .... import ops
from .... import autograd as ag

Tensor = torch.Tensor

# Philox stream position: (torch.initial_seed(), call counter) -> reproducible
# under torch.manual_seed for a fixed sequence of calls.
_RNG_CALLS = itertools.count()


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise ops.L2BError('l2hmc_b200 needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def checkSU(x: Tensor) -> tuple[Tensor, Tensor]:
    """utils.py:376-391 -> (avg, max) per chain"""
    if x.dim() == 8:
        return ops.su3_check(x)
    # generic batch: treat dim 0 as the chain axis of a 1x1x1xN lattice of links
    nb = x.shape[0]
    n = x[0].numel() // 9
    pad = (-n) % 4
    if pad:
        eye = torch.eye(3, dtype=x.dtype, device=x.device).expand(nb, pad, 3, 3)
        x = torch.cat([x.reshape(nb, n, 3, 3), eye], 1)
        n += pad
    # padding with identities adds zero deviation; rescale the mean accordingly
    avg, mx = ops.su3_check(x.reshape(nb, 4, 1, 1, 1, n // 4, 3, 3))
    if pad:
        avg = avg * (n / (n - pad)) ** 0.5
    return avg, mx


def projectSU(x: Tensor) -> Tensor:
    return ops.su3_project(x)


def projectTAH(x: Tensor) -> Tensor:
    return ops.su3_tah(x)


def su3_to_vec(x: Tensor) -> Tensor:
    return ops.su3_to_vec(x)


def vec_to_su3(v: Tensor) -> Tensor:
    return ops.su3_from_vec(v)


def norm2(x: Tensor, axis: Sequence[int] = (-2, -1), exclude: Optional[Sequence[int]] = None) -> Tensor:
    """utils.py:157-168 (plain reduction; not on the integrator's inner loop)"""
    if x.is_complex():
        x = x.abs()
    n = x.square()
    if exclude is None:
        return n if len(axis) == 0 else n.sum(tuple(axis))
    return n.sum([i for i in range(len(n.shape)) if i not in exclude])


def eyeOf(x: Tensor) -> Tensor:
    """broadcastable real identity (utils.py:134-141)"""
    eye = torch.zeros([1] * (x.dim() - 2) + [*x.shape[-2:]], device=x.device)
    eye[-2:] = torch.eye(x.shape[-1], device=x.device)
    return eye


def checkU(x: Tensor) -> tuple[Tensor, Tensor]:
    """average / maximum deviation of X^+X from 1 (utils.py:362-374); a diagnostic, plain torch"""
    nc = x.shape[-1]
    d = norm2(x.mH @ x - eyeOf(x)).flatten(1)
    c = 2 * (nc * nc + 1)
    return (d.mean(-1) / c).sqrt(), (d.max(-1)[0] / c).sqrt()


def rsqrtPHM3f(tr: Tensor, p2: Tensor, det: Tensor) -> tuple[Tensor, Tensor, Tensor]:
    """coefficients of X^{-1/2} = c0 + c1 X + c2 X^2 for positive Hermitian 3x3 X from its
    invariants (utils.py:227-317, clamps included) -- the same closed form `k_unary<PROJECT>` uses
    per link; this torch version serves callers outside the integrator"""
    tr3, p23 = tr / 3.0, p2 / 3.0
    tr32 = tr3 * tr3
    q = (0.5 * (p23 - tr32)).abs()
    r = 0.25 * tr3 * (5.0 * tr32 - p2) - 0.5 * det
    sq = q.sqrt()
    isq3 = (1.0 / (q * sq)).clamp(-3e38, 3e38)
    rsq3 = (r * isq3).clamp(-1.0, 1.0).clamp(-1.0 + 1e-12, 1.0 - 1e-12)
    t = torch.acos(rsq3) / 3.0
    sqc, sqs = sq * t.cos(), (3.0 ** 0.5) * sq * t.sin()
    ll = tr3 + sqc
    e0, e1, e2 = tr3 - 2.0 * sqc, ll + sqs, ll - sqs
    s0, s1, s2 = e0.abs().sqrt(), e1.abs().sqrt(), e2.abs().sqrt()
    u, w = s0 + s1 + s2, s0 * s1 * s2
    di = 1.0 / (w * (s0 + s1) * (s0 + s2) * (s1 + s2))
    c0 = di * (w * u * u + e0 * s0 * (e1 + e2) + e1 * s1 * (e0 + e2) + e2 * s2 * (e0 + e1))
    return c0, -(tr * u + w) * di, u * di


def rsqrtPHM3(x: Tensor) -> Tensor:
    """X^{-1/2} of a positive Hermitian 3x3 (utils.py:320-330)"""
    tr = torch.diagonal(x, dim1=-2, dim2=-1).sum(-1).real
    x2 = x @ x
    p2 = torch.diagonal(x2, dim1=-2, dim2=-1).sum(-1).real
    c0, c1, c2 = (c[..., None, None].to(x.dtype) for c in rsqrtPHM3f(tr, p2, torch.linalg.det(x).real))
    return c0 * torch.eye(3, dtype=x.dtype, device=x.device) + c1 * x + c2 * x2


def projectU(x: Tensor) -> Tensor:
    """X (X^+X)^{-1/2}: the unitary polar factor without the determinant phase (utils.py:333-338).
    projectSU(X) = projectU(X) e^{-i arg det / 3} and arg det projectU(X) = arg det X, so the
    kernel's projectSU is re-phased instead of repeating the eigen-solve in torch."""
    y = ops.su3_project(x)
    ph = torch.angle(torch.linalg.det(x)) / 3.0
    return y * torch.polar(torch.ones_like(ph), ph)[..., None, None]


def randTAH3(shape: Sequence[int], device=None) -> Tensor:
    """utils.py:171-195; `shape` is the batch shape [nb, 4, T, X, Y, Z]"""
    shape = tuple(int(s) for s in shape)
    device = _device() if device is None else torch.device(device)
    if len(shape) == 6 and shape[1] == 4:
        nb, dims = shape[0], shape[2:]
    else:  # arbitrary batch: generate as one chain and reshape
        n = 1
        for s in shape:
            n *= s
        pad = (-n) % 4
        p = ops.su3_rand_momentum(1, [1, 1, 1, (n + pad) // 4], torch.initial_seed(), next(_RNG_CALLS), device)
        return p.reshape(-1, 3, 3)[:n].reshape(*shape, 3, 3)
    if torch.cuda.is_current_stream_capturing():
        # inside a CUDA-graph capture the host call counter would be frozen into the graph:
        # use a device-resident counter in a disjoint offset range (bit 62 set) instead
        return ops.su3_rand_momentum(nb, dims, torch.initial_seed(), (1 << 62) + (next(_RNG_CALLS) << 32), device,
                                     offset_dev=_graph_counter(device))
    return ops.su3_rand_momentum(nb, dims, torch.initial_seed(), next(_RNG_CALLS), device)


_GRAPH_COUNTERS: dict = {}


def _graph_counter(device) -> Tensor:
    """one int64 device scalar per device, allocated OUTSIDE any capture (first use warms it up)"""
    key = torch.device(device).index
    if key not in _GRAPH_COUNTERS:
        if torch.cuda.is_current_stream_capturing():
            raise ops.L2BError('run one SU(3) step eagerly before capturing (the RNG counter is not allocated yet)')
        _GRAPH_COUNTERS[key] = torch.zeros(1, dtype=torch.int64, device=device)
    return _GRAPH_COUNTERS[key]


class SU3(Group):
    def __init__(self) -> None:
        super().__init__(dim=4, shape=[3, 3], dtype=torch.complex128, name='SU3')

    # -- integrator ops ----------------------------------------------------
    def update_gauge(self, x: Tensor, p: Tensor) -> Tensor:
        """exp(p) @ x   (group.py:45-50)"""
        if x.dim() == 8:
            return ops.su3_update_gauge(x, p, 1.0)
        return ops.su3_exp(p) @ x

    def exp(self, x: Tensor) -> Tensor:
        return ops.su3_exp(x)

    def projectTAH(self, x: Tensor) -> Tensor:
        return ops.su3_tah(x)

    def projectSU(self, x: Tensor) -> Tensor:
        return ag.SU3Project.apply(x)

    def compat_proj(self, x: Tensor) -> Tensor:
        return ag.SU3Project.apply(x)

    def compat_proju(self, u: Tensor, x: Tensor) -> Tensor:
        """group.py:149-165: solve u A = x for the FIRST matrix of the batch only (the reference indexes the
        solution with [0] and lets the rest broadcast), then the traceless anti-Hermitian part of A.
        Off the integrator path (nothing in the reference calls it): plain torch."""
        n = x.shape[-1]
        a = torch.linalg.solve(u, x)[0]
        b = (a - a.mH) / 2.
        tr = torch.einsum('...ii->...', b)
        b = b - (tr / n)[..., None, None] * torch.eye(n, dtype=b.dtype, device=b.device)
        assert torch.abs(torch.einsum('...ii->...', b).mean()) < 1e-6
        return b

    def kinetic_energy(self, p: Tensor) -> Tensor:
        """0.5 * sum(|P|_F^2 - 8)   (group.py:125-126)"""
        return ag.SU3Kinetic.apply(p)

    def group_to_vec(self, x: Tensor, dtype: torch.dtype = torch.float64) -> Tensor:
        """su3_to_vec(projectSU(x)) in one kernel   (group.py:138-147); `dtype` lets the
        caller receive the 8 reals per link directly in the nets' element type"""
        return ag.SU3GroupToVec.apply(x, dtype)

    def vec_to_group(self, x: Tensor) -> Tensor:
        """projectSU(vec_to_su3(x))   (group.py:128-136)"""
        return ops.su3_project(ops.su3_from_vec(x))

    def random(self, shape: Sequence[int]) -> Tensor:
        """projectSU(N(0,1) + i N(0,1))   (group.py:113-119)"""
        dev = _device()
        r = torch.randn(*shape, dtype=torch.float64, device=dev)
        i = torch.randn(*shape, dtype=torch.float64, device=dev)
        return ops.su3_project(torch.complex(r, i))

    def random_momentum(self, shape: Sequence[int]) -> Tensor:
        """randTAH3(shape[:-2])   (group.py:121-123)"""
        return randTAH3(list(shape)[:-2])

    def checkU(self, x: Tensor) -> tuple[Tensor, Tensor]:
        return checkU(x)

    def projectU(self, x: Tensor) -> Tensor:
        return projectU(x)

    def rsqrtPHM3(self, x: Tensor) -> Tensor:
        return rsqrtPHM3(x)

    def rsqrtPHM3f(self, tr: Tensor, p2: Tensor, det: Tensor):
        return rsqrtPHM3f(tr, p2, det)

    def diff_trace(self, x: Tensor) -> Tensor:
        """a TODO stub in the reference too (group.py:80-86)"""
        return x

    def diff2trace(self, x: Tensor) -> Tensor:
        return x

    def checkSU(self, x: Tensor) -> tuple[Tensor, Tensor]:
        return checkSU(x)

    # -- thin algebra helpers kept for API parity (the lattice never calls these:
    #    its products are fused inside the stencil kernels) --------------------
    def mul(self, a: Tensor, b: Tensor, adjoint_a: bool = False, adjoint_b: bool = False) -> Tensor:
        if adjoint_a:
            a = a.adjoint()
        if adjoint_b:
            b = b.adjoint()
        return a @ b

    def adjoint(self, x: Tensor) -> Tensor:
        return x.adjoint()

    def trace(self, x: Tensor) -> Tensor:
        return torch.diagonal(x, dim1=-2, dim2=-1).sum(-1)

    def norm2(self, x: Tensor, axis: Sequence[int] = (-2, -1), exclude=None) -> Tensor:
        return norm2(x, axis, exclude)

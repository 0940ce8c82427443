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

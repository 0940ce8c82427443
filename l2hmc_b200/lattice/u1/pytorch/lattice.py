"""`LatticeU1` with the reference's method surface
(`lattice/u1/pytorch/lattice.py:50-317`), backed by libl2b's U(1) kernels.
`x`: `[nb, 2, T, X]` real angles on CUDA (or flattened `[nb, -1]`)."""
from __future__ import annotations

from math import pi as PI
from typing import Optional

import torch
from torch.special import i0, i1

from ....configs import Charges, LatticeMetrics
from ....group.u1.pytorch.group import U1Phase
from ....lattice.lattice import Lattice
from .... import ops
from .... import autograd as ag

TWOPI = 2. * PI
Tensor = torch.Tensor


def _f(beta) -> float:
    return float(beta.detach()) if isinstance(beta, torch.Tensor) else float(beta)


def plaq_exact(beta):
    """I1(beta)/I0(beta)   (lattice.py:37-42)"""
    beta = torch.as_tensor(beta, dtype=torch.float32)
    return (i1(beta) / i0(beta)).to(torch.get_default_dtype())


def area_law(beta: float, nplaqs: int):
    beta = torch.as_tensor(beta)
    return (i1(beta) / i0(beta)) ** nplaqs


def project_angle(x: Tensor) -> Tensor:
    return x - TWOPI * torch.floor((x + PI) / TWOPI)


class LatticeU1(Lattice):
    def __init__(self, nchains: int, shape: list[int]):
        assert len(shape) == 2
        self.g = U1Phase()
        self.nt, self.nx = shape
        self.nplaqs = self.nt * self.nx
        super().__init__(group=self.g, nchains=nchains, shape=list(shape))

    def _field(self, x: Tensor) -> Tensor:
        if x.dim() != 4:
            x = x.reshape(x.shape[0], *self.xshape)
        return x

    # -- kernels ---------------------------------------------------------------
    def wilson_loops(self, x: Tensor) -> Tensor:
        """differentiable (the loss back-propagates through it, loss/pytorch/loss.py:194-197)"""
        return ag.U1WilsonLoops.apply(x, list(self._lattice_shape))

    def _obs(self, x: Tensor, beta=1.0) -> Tensor:
        """[nb, 4] = (action, plaq, sinQ, intQ) in one pass"""
        return ops.u1_observables(self._field(x.detach()), _f(beta))

    def action(self, x: Tensor, beta) -> Tensor:
        """differentiable: backward is the force kernel"""
        return ag.U1Action.apply(x, _f(beta), list(self._lattice_shape))

    def grad_action(self, x: Tensor, beta, create_graph: bool = True) -> Tensor:
        """analytic dS/dx (reference: autograd, lattice.py:102-117), shaped like x.
        Differentiable once more, like the reference's create_graph=True: the
        backward is the Hessian-vector-product kernel."""
        if not create_graph:
            x = x.detach()
        return ag.U1Force.apply(x, _f(beta), list(self._lattice_shape))

    def action_with_grad(self, x: Tensor, beta) -> tuple[Tensor, Tensor]:
        return self.action(x, beta), self.grad_action(x, beta)

    def kinetic_energy(self, v: Tensor) -> Tensor:
        return self.g.kinetic_energy(v)

    # -- wloops-based API kept for the loss / trainer (lattice.py:188-228) --------
    def _action(self, wloops: Tensor, beta) -> Tensor:
        return _f(beta) * (1. - wloops.cos()).sum((1, 2))

    def _plaqs(self, wloops: Tensor) -> Tensor:
        return wloops.cos().mean((1, 2))

    def plaqs(self, x: Optional[Tensor] = None, wloops: Optional[Tensor] = None) -> Tensor:
        if wloops is None:
            if x is None:
                raise ValueError('One of `x` or `wloops` must be specified.')
            return self._obs(x)[:, 1]
        return self._plaqs(wloops)

    def _sin_charges(self, wloops: Tensor) -> Tensor:
        return wloops.sin().sum((1, 2)) / TWOPI

    def _int_charges(self, wloops: Tensor) -> Tensor:
        return project_angle(wloops).sum((1, 2)) / TWOPI

    def _charges(self, wloops: Tensor) -> Charges:
        return Charges(intQ=self._int_charges(wloops), sinQ=self._sin_charges(wloops))

    def charges(self, x: Optional[Tensor] = None, wloops: Optional[Tensor] = None) -> Charges:
        if wloops is not None:
            return self._charges(wloops)
        o = self._obs(x)
        return Charges(intQ=o[:, 3], sinQ=o[:, 2])

    def sin_charges(self, x: Optional[Tensor] = None, wloops: Optional[Tensor] = None) -> Tensor:
        return self.charges(x, wloops).sinQ

    def int_charges(self, x: Optional[Tensor] = None, wloops: Optional[Tensor] = None) -> Tensor:
        return self.charges(x, wloops).intQ

    def calc_metrics(self, x: Tensor) -> dict:
        o = self._obs(x)
        return {'plaqs': o[:, 1], 'intQ': o[:, 3], 'sinQ': o[:, 2]}

    def plaqs_diff(self, beta: float, x: Optional[Tensor] = None, wloops: Optional[Tensor] = None) -> Tensor:
        plaqs = self.plaqs(x=x, wloops=wloops)
        return plaq_exact(beta).to(plaqs.device) * torch.ones_like(plaqs) - plaqs

    def wilson_loops4x4(self, x: Tensor) -> Tensor:
        """4x4 loops (lattice.py:161-186): `l2b_u1_wilson_loops4x4`; under autograd the same sixteen rolled terms
        as torch ops"""
        x = self._field(x)
        if x.is_cuda and not (torch.is_grad_enabled() and x.requires_grad):
            return ops.u1_wilson_loops4x4(x.detach()).permute(2, 1, 0)        # upstream's `.T` of a 3-D tensor
        xu, xv = x[:, 0], x[:, 1]
        return (
            xu + xu.roll(-1, dims=2) + xu.roll(-2, dims=2) + xu.roll(-3, dims=2) + xu.roll(-4, dims=2)
            + xv.roll((-4, -1), dims=(2, 1)) + xv.roll((-4, -2), dims=(2, 1)) + xv.roll((-4, -3), dims=(2, 1))
            - xu.roll((-3, -4), dims=(2, 1)) - xu.roll((-2, -4), dims=(2, 1)) - xu.roll((-1, -4), dims=(2, 1))
            - xv.roll(-4, dims=1) - xv.roll(-3, dims=1) - xv.roll(-2, dims=1) - xv.roll(-1, dims=1) - xv
        ).T

    def _plaqs4x4(self, wloops4x4: Tensor) -> Tensor:
        """lattice.py:205-206"""
        return wloops4x4.cos().mean((1, 2))

    def plaqs4x4(self, x: Optional[Tensor] = None, wloops4x4: Optional[Tensor] = None) -> Tensor:
        if wloops4x4 is None:
            if x is None:
                raise ValueError('One of `x` or `wloops` must be specified.')
            wloops4x4 = self.wilson_loops4x4(x)
        return self._plaqs4x4(wloops4x4)

    def _get_wloops(self, x: Optional[Tensor] = None) -> Tensor:
        """lattice.py:230-236"""
        if x is None:
            raise ValueError('One of `x` or `wloops` must be specified.')
        return self.wilson_loops(x)

    def draw_uniform_batch(self, requires_grad: bool = True) -> Tensor:
        """a batch of configurations uniform in (-pi, pi)   (lattice.py:67-71), on the GPU"""
        return self.g.random(list(self._shape)).detach().requires_grad_(requires_grad)

    def plaq_loss(self, acc: Tensor, x1: Optional[Tensor] = None, x2: Optional[Tensor] = None,
                  wl1: Optional[Tensor] = None, wl2: Optional[Tensor] = None) -> Tensor:
        """-mean_b[ acc * sum 2 (1 - cos(w2 - w1)) + 1e-4 ]   (lattice.py:278-292)"""
        w1 = self._get_wloops(x1) if wl1 is None else wl1
        w2 = self._get_wloops(x2) if wl2 is None else wl2
        return -(acc * (2. * (1. - (w2 - w1).cos())).sum((1, 2)) + 1e-4).mean(0)

    def charge_loss(self, acc: Tensor, x1: Optional[Tensor] = None, x2: Optional[Tensor] = None,
                    wl1: Optional[Tensor] = None, wl2: Optional[Tensor] = None) -> Tensor:
        """-mean_b[ acc * (sinQ2 - sinQ1)^2 + 1e-4 ]   (lattice.py:294-308)"""
        w1 = self._get_wloops(x1) if wl1 is None else wl1
        w2 = self._get_wloops(x2) if wl2 is None else wl2
        return -(acc * (self._sin_charges(w2) - self._sin_charges(w1)) ** 2 + 1e-4).mean(0)

    def observables(self, x: Tensor) -> LatticeMetrics:
        wloops = self.wilson_loops(x)
        return LatticeMetrics(p4x4=self.plaqs4x4(x=x), plaqs=self.plaqs(wloops=wloops),
                              charges=self.charges(wloops=wloops))

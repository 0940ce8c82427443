"""`LatticeSU3` with the reference's method surface
(`lattice/su3/pytorch/lattice.py:41-349`), backed by the stencil kernels of
libl2b: plaquette traces / per-chain sums (`l2b_su3_wilson_loops`,
`l2b_su3_plaq_sums`) and the analytic staple force (`l2b_su3_force`) in place of
the reference's autograd-of-the-action.

`x`: `[nb, 4, T, X, Y, Z, 3, 3]` complex128 on CUDA (or flattened `[nb, -1]`).
"""
from __future__ import annotations

import os
from typing import Optional

import numpy as np
import torch

from ....configs import Charges
from ....group.su3.pytorch.group import SU3
from ....lattice.lattice import Lattice
from .... import ops
from .... import autograd as ag

Tensor = torch.Tensor


def pbc(tup, shape) -> list:
    """site coordinates wrapped into the lattice (lattice.py:33-34)"""
    return np.mod(tup, shape).tolist()


def mat_adj(mat):
    """conjugate transpose of a single matrix (lattice.py:37-38)"""
    return mat.conj().T


def _f(beta) -> float:
    return float(beta.detach()) if isinstance(beta, torch.Tensor) else float(beta)


class LatticeSU3(Lattice):
    """4D lattice with SU(3) links."""
    dim = 4

    def __init__(self, nchains: int, shape: list[int], c1: float = 0.0) -> None:
        assert len(shape) == 4
        self.g = SU3()
        self.nt, self.nx, self.ny, self.nz = shape
        self.c1 = float(c1)      # c1 != 0 (DBW2 / rectangle term): plaquette part on the kernels, rectangle
        # `rect_kernel`: evaluate the improved action / force with the hand-written rectangle-staple kernel
        # (l2b_su3_force_c1); L2B_RECT_KERNEL=0 forces the ATen path everywhere.
        self.rect_kernel = os.environ.get('L2B_RECT_KERNEL', '1') == '1'
        # under autograd too (adjoint kernel l2b_su3_action_grad_c1; body pinned on torch autograd by
        # tests/test_hostemu.py, launch by tests/test_gpu_su3.py::test_rectangle_kernel_gradients against the
        # reference's autograd goldens).  L2B_RECT_KERNEL_AUTOGRAD=0: rectangle part as ATen ops under autograd.
        self.rect_kernel_autograd = os.environ.get('L2B_RECT_KERNEL_AUTOGRAD', '1') == '1'
        super().__init__(group=self.g, nchains=nchains, shape=list(shape))

    def _field(self, x: Tensor) -> Tensor:
        if x.dim() != 8:
            x = x.reshape(x.shape[0], *self._shape[1:])
        return x

    def coeffs(self, beta) -> dict:
        b = _f(beta)
        return {'plaq': b * (1.0 - 8.0 * self.c1), 'rect': b * self.c1}

    # -- observables ---------------------------------------------------------
    def _sums(self, x: Tensor) -> Tensor:
        """[nb, 2] = (sum Re tr P, sum Im tr P) in one pass over the links"""
        return ops.su3_plaq_sums(self._field(x.detach()))

    def wilson_loops(self, x: Tensor) -> Tensor:
        """ps[6, nb, T, X, Y, Z]   (lattice.py:157-199,242-244); differentiable"""
        return ag.SU3WilsonLoops.apply(self._field(x))

    # -- matrix-valued loop fields: debugging / analysis API of the reference, off the integrator path ---------
    def _link_staple_op(self, link: Tensor, staple: Tensor) -> Tensor:
        """lattice.py:93-94"""
        return self.g.mul(link, staple)

    def _plaquette(self, x: Tensor, u: int, v: int) -> Tensor:
        """U_u(n) U_v(n+u) U_u(n+v)^+ U_v(n)^+ as a matrix field [nb, T, X, Y, Z, 3, 3] (lattice.py:114-126)"""
        x = self._field(x)
        xu, xv = x[:, u], x[:, v]
        return (xu @ xv.roll(-1, dims=u + 1)) @ (xv @ xu.roll(-1, dims=v + 1)).mH

    def _trace_plaquette(self, x: Tensor, u: int, v: int) -> Tensor:
        """lattice.py:128-130"""
        return self.g.trace(self._plaquette(x, u, v))

    def _rectangles(self, x: Tensor, u: int, v: int) -> tuple[Tensor, Tensor]:
        """the 2x1 (two steps along u) and 1x2 rectangle matrices of plane (u, v), built from the four
        three-link staples of the plaquette exactly as the reference does (lattice.py:96-112)"""
        x = self._field(x)
        xu, xv = x[:, u], x[:, v]
        su, sv = u + 1, v + 1                           # tensor axes of the two directions
        fwd_uv = xu @ xv.roll(-1, dims=su)              # U_u(n) U_v(n+u)
        fwd_vu = xv @ xu.roll(-1, dims=sv)              # U_v(n) U_u(n+v)
        open_v = xv.mH @ fwd_uv                         # U_v(n)^+ U_u(n) U_v(n+u)
        open_u = xu.mH @ fwd_vu                         # U_u(n)^+ U_v(n) U_u(n+v)
        cap_u = fwd_uv @ xu.roll(-1, dims=sv).mH        # U_u(n) U_v(n+u) U_u(n+v)^+
        cap_v = fwd_vu @ xv.roll(-1, dims=su).mH        # U_v(n) U_u(n+v) U_v(n+u)^+
        return open_u @ cap_u.roll(-1, dims=su).mH, open_v @ cap_v.roll(-1, dims=sv).mH

    def _plaquette_field(self, x: Tensor, needs_rect: bool = False) -> tuple[list, list]:
        """six plaquette matrix fields (planes u > v in the reference's order) and the twelve rectangle
        fields, zeros unless `needs_rect` (lattice.py:132-156)"""
        x = self._field(x)
        plaqs, rects = [], []
        for u in range(1, self.dim):
            for v in range(u):
                plaq = self._plaquette(x, u, v)
                plaqs.append(plaq)
                rects.extend(self._rectangles(x, u, v) if needs_rect
                             else (torch.zeros_like(plaq), torch.zeros_like(plaq)))
        return plaqs, rects

    def _rect_traces(self, x: Tensor) -> Tensor:
        """traces of the 2x1 and 1x2 rectangles, rs[12, nb, T, X, Y, Z] (lattice.py:180-196), as ATen ops:
        the differentiable form used under autograd; HMC / eval use the rectangle-staple kernel
        (`l2b_su3_force_c1`)"""
        x = self._field(x)
        rs = []
        for u in range(1, 4):
            for v in range(0, u):
                r21, r12 = self._rectangles(x, u, v)
                rs.append(self.g.trace(r21))
                rs.append(self.g.trace(r12))
        return torch.stack(rs)

    def plaq_loss(self, acc: Tensor, x1: Optional[Tensor] = None, x2: Optional[Tensor] = None,
                  wloops1: Optional[Tensor] = None, wloops2: Optional[Tensor] = None):
        """a TODO stub in the reference (lattice.py:351-359): nothing is computed there either; the SU(3) loss
        terms live in `LatticeLoss` (loss/pytorch/loss.py)"""
        return None

    def charge_loss(self, acc: Tensor, x1: Optional[Tensor] = None, x2: Optional[Tensor] = None,
                    wloops1: Optional[Tensor] = None, wloops2: Optional[Tensor] = None):
        """a TODO stub in the reference (lattice.py:361-369)"""
        return None

    def _rect_action(self, x: Tensor, beta) -> Tensor:
        """-(beta c1 / 3) sum Re tr R"""
        rs = self._rect_traces(x)
        return rs.real.sum(tuple(range(2, rs.dim()))).sum(0) * (-self.coeffs(beta)['rect'] / 3.0)

    def _wilson_loops(self, x: Tensor, needs_rect: bool = False) -> tuple[Tensor, Tensor]:
        ps = self.wilson_loops(x)
        if needs_rect:
            return ps, self._rect_traces(x)
        return ps, torch.zeros((12, *ps.shape[1:]), dtype=ps.dtype, device=ps.device)

    def action(self, x: Tensor, beta) -> Tensor:
        """S = -(beta (1 - 8 c1) / 3) sum Re tr P - (beta c1 / 3) sum Re tr R   (lattice.py:252-269);
        differentiable (plaquette part: adjoint = the staple-sum kernel)"""
        if self._use_rect_kernel(x):
            sums = ops.su3_force_c1(self._field(x), _f(beta), self.c1, want_force=False, want_sums=True)
            return (self.coeffs(beta)['plaq'] * sums[:, 0] + self.coeffs(beta)['rect'] * sums[:, 1]) * (-1.0 / 3.0)
        if self.c1 != 0.0 and self.rect_kernel and self.rect_kernel_autograd:
            return ag.SU3ActionC1.apply(self._field(x), _f(beta), self.c1)
        s = ag.SU3Action.apply(self._field(x), self.coeffs(beta)['plaq'])
        if self.c1 != 0.0:
            s = s + self._rect_action(x, beta)
        return s

    def _action(self, wloops, beta) -> Tensor:
        """NB: the reference's `_action` has no minus sign (lattice.py:271-285)"""
        ps = wloops[0] if isinstance(wloops, (tuple, list)) else wloops
        psum = ps.real.sum(tuple(range(2, ps.dim()))).sum(0)
        act = self.coeffs(beta)['plaq'] * psum
        if self.c1 != 0.0 and isinstance(wloops, (tuple, list)):
            rs = wloops[1]
            act = act + self.coeffs(beta)['rect'] * rs.real.sum(tuple(range(2, rs.dim()))).sum(0)
        return act / 3.0

    def _plaquettes(self, x: Tensor) -> Tensor:
        return self._sums(x)[:, 0] / (6 * 3 * self.volume)

    def _plaqs(self, wloops: Tensor) -> Tensor:
        psum = wloops.real.sum(tuple(range(2, wloops.dim()))).sum(0)
        return psum / (6 * 3 * self.volume)

    def plaqs(self, x: Optional[Tensor] = None, wloops: Optional[Tensor] = None) -> Tensor:
        if wloops is None:
            assert x is not None
            return self._plaquettes(x)
        return self._plaqs(wloops)

    def _charges(self, wloops: Tensor) -> Charges:
        qsum = wloops.imag.sum(tuple(range(2, wloops.dim()))).sum(0)
        return Charges(intQ=qsum / (32 * (np.pi ** 2)), sinQ=qsum / (6 * 3 * self.volume))

    def _int_charges(self, wloops: Tensor) -> Tensor:
        return self._charges(wloops).intQ

    def _sin_charges(self, wloops: Tensor) -> Tensor:
        return self._charges(wloops).sinQ

    def charges(self, x: Optional[Tensor] = None, wloops: Optional[Tensor] = None) -> Charges:
        if wloops is not None:
            return self._charges(wloops)
        q = self._sums(x)[:, 1]
        return Charges(intQ=q / (32 * (np.pi ** 2)), sinQ=q / (6 * 3 * self.volume))

    def int_charges(self, x: Optional[Tensor] = None, wloops: Optional[Tensor] = None) -> Tensor:
        return self.charges(x, wloops).intQ

    def sin_charges(self, x: Optional[Tensor] = None, wloops: Optional[Tensor] = None) -> Tensor:
        return self.charges(x, wloops).sinQ

    # -- energies / force ----------------------------------------------------
    def kinetic_energy(self, v: Tensor) -> Tensor:
        return self.g.kinetic_energy(self._field(v))

    def grad_action(self, x: Tensor, beta) -> Tensor:
        """(beta/3) TAH(U A), analytic; equals the reference's
        projectTAH(autograd(S) @ x^+) (lattice.py:299-308).  Like the reference
        (no create_graph) the result is a constant w.r.t. later backprop."""
        if self._use_rect_kernel(x):
            return ops.su3_force_c1(self._field(x), _f(beta), self.c1)
        if self.c1 != 0.0 and self.rect_kernel and self.rect_kernel_autograd:
            return ag.SU3ForceC1.apply(self._field(x), _f(beta), self.c1)
        f = ag.SU3Force.apply(self._field(x), self.coeffs(beta)['plaq'])
        if self.c1 != 0.0:
            f = f + self._rect_force(self._field(x), beta)
        return f

    def _use_rect_kernel(self, x: Tensor) -> bool:
        """the rectangle kernel has no adjoint: only where nothing will be back-propagated"""
        return (self.c1 != 0.0 and self.rect_kernel
                and not (torch.is_grad_enabled() and (x.requires_grad or isinstance(self.c1, Tensor))))

    def _rect_force(self, x: Tensor, beta) -> Tensor:
        """projectTAH(dS_rect/dx @ x^+) with dS_rect/dx from ATen autograd, detached like the
        reference's `dsdx` (lattice.py:299-308); the explicit `@ x^+` stays differentiable"""
        with torch.enable_grad():
            xr = x.detach().requires_grad_(True)
            dsdx, = torch.autograd.grad(self._rect_action(xr, beta).sum(), xr)
        y = dsdx.detach() @ x.mH
        a = 0.5 * (y - y.mH)
        return a - torch.diagonal(a, dim1=-2, dim2=-1).sum(-1)[..., None, None] / 3.0 * torch.eye(
            3, dtype=a.dtype, device=a.device)

    def action_with_grad(self, x: Tensor, beta) -> tuple[Tensor, Tensor]:
        """one force pass yields both (lattice.py:287-297)"""
        if self._use_rect_kernel(x.detach()):
            f, sums = ops.su3_force_c1(self._field(x.detach()), _f(beta), self.c1, want_sums=True)
            return (self.coeffs(beta)['plaq'] * sums[:, 0] + self.coeffs(beta)['rect'] * sums[:, 1]) * (-1.0 / 3.0), f
        if self.c1 != 0.0:
            return self.action(x, beta).detach(), self.grad_action(x, beta).detach()
        f, ps = ops.su3_force(self._field(x.detach()), _f(beta), want_plaq_sum=True)
        return ps * (-_f(beta) / 3.0), f

    def calc_metrics(self, x: Tensor, beta=None, xinit: Optional[Tensor] = None) -> dict:
        """lattice.py:310-349"""
        sums = self._sums(x)
        V18 = 6 * 3 * self.volume
        plaqs = sums[:, 0] / V18
        intQ, sinQ = sums[:, 1] / (32 * np.pi ** 2), sums[:, 1] / V18
        metrics = {'plaqs': plaqs, 'sinQ': sinQ, 'intQ': intQ}
        if beta is not None:
            s, dsdx = self.action_with_grad(x, beta)
            metrics.update({'action': s, 'dsdx': dsdx})
            if xinit is not None:
                s_, dsdx_ = self.action_with_grad(xinit, beta)
                metrics.update({'daction': (s - s_).abs(), 'dsdx': (dsdx - dsdx_).abs()})
        if xinit is not None:
            sums_ = self._sums(xinit)
            metrics.update({
                'dplaqs': (plaqs - sums_[:, 0] / V18).abs(),
                'dQint': (intQ - sums_[:, 1] / (32 * np.pi ** 2)).abs(),
                'dQsin': (sinQ - sums_[:, 1] / V18).abs(),
            })
        return metrics

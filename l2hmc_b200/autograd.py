"""torch.autograd.Function wrappers that give the libl2b kernels a backward, so
that L2HMC training (`Trainer.train_step`: loss.backward() through
`Dynamics.forward`, trainers/pytorch/trainer.py:1316-1367) runs on the CUDA path.
The reference obtains every derivative from autograd over ATen ops; here each
forward kernel has a hand-written adjoint kernel (include/l2b.h, "adjoints").

U(1) is complete (force is differentiated to second order, like the reference's
`create_graph=True`, lattice/u1/pytorch/lattice.py:106).  Step sizes enter as
0-dim tensors so their gradient flows back to the `xeps`/`veps` parameters.
"""
from __future__ import annotations

import torch

from . import ops

Tensor = torch.Tensor


def _eps_grad(geps_per_chain: Tensor, eps: Tensor):
    return geps_per_chain.sum().to(eps.dtype).reshape(eps.shape)


class U1WilsonLoops(torch.autograd.Function):
    """w = x0 + roll(x1,-1,T) - roll(x0,-1,X) - x1"""

    @staticmethod
    def forward(ctx, x, shape):
        ctx.xshape = x.shape
        return ops.u1_wilson_loops(x.detach(), shape)

    @staticmethod
    def backward(ctx, gw):
        return ops.u1_wilson_loops_bwd(gw).reshape(ctx.xshape), None


class U1Action(torch.autograd.Function):
    """S[b] = beta sum(1 - cos w);  dS/dx = force"""

    @staticmethod
    def forward(ctx, x, beta, shape):
        ctx.save_for_backward(x)
        ctx.beta, ctx.shape = beta, shape
        return ops.u1_observables(x.detach(), beta, shape)[:, 0]

    @staticmethod
    def backward(ctx, gs):
        x, = ctx.saved_tensors
        f = ops.u1_force(x.detach(), ctx.beta, ctx.shape)
        return ops.rowscale(f, gs).reshape(x.shape), None, None


class U1Force(torch.autograd.Function):
    """F = dS/dx; backward = Hessian-vector product (second derivative of S)"""

    @staticmethod
    def forward(ctx, x, beta, shape):
        ctx.save_for_backward(x)
        ctx.beta, ctx.shape = beta, shape
        return ops.u1_force(x.detach(), beta, shape).reshape(x.shape)

    @staticmethod
    def backward(ctx, gf):
        x, = ctx.saved_tensors
        return ops.u1_force_bwd(x.detach(), ctx.beta, gf, ctx.shape).reshape(x.shape), None, None


class U1Kinetic(torch.autograd.Function):
    @staticmethod
    def forward(ctx, v):
        ctx.save_for_backward(v)
        return ops.u1_kinetic(v.detach())

    @staticmethod
    def backward(ctx, g):
        v, = ctx.saved_tensors
        return ops.rowscale(v.detach(), g)


class U1CompatProj(torch.autograd.Function):
    """((x + pi) mod 2 pi) - pi: unit derivative almost everywhere"""

    @staticmethod
    def forward(ctx, x):
        return ops.u1_compat_proj(x.detach())

    @staticmethod
    def backward(ctx, g):
        return g


def _opt(a):
    return None if a is None else a.detach()


class U1VUpdate(torch.autograd.Function):
    """(v', logdet) = vupdate(v, F, s, t, q; eps, sign)   (dynamics.py:1266-1297)"""

    @staticmethod
    def forward(ctx, v, force, s, t, q, eps, sign):
        ctx.save_for_backward(v, force, s, t, q, eps)
        ctx.sign = sign
        out, logdet = ops.u1_vupdate(v.detach(), force.detach(), _opt(s), _opt(t), _opt(q), float(eps), sign)
        return out, logdet

    @staticmethod
    def backward(ctx, gout, glogdet):
        v, force, s, t, q, eps = ctx.saved_tensors
        gv, gf, gs, gt, gq, geps = ops.u1_vupdate_bwd(v.detach(), force.detach(), _opt(s), _opt(t), _opt(q), float(eps),
                                                      ctx.sign, gout, glogdet)
        rs = lambda g, ref: None if (g is None or ref is None) else g.reshape(ref.shape).to(ref.dtype)  # noqa: E731
        return (gv.reshape(v.shape), gf.reshape(force.shape).to(force.dtype), rs(gs, s), rs(gt, t), rs(gq, q),
                _eps_grad(geps, eps), None)


class U1XUpdate(torch.autograd.Function):
    """(x', logdet) = xupdate(x, v, s, t, q; mask, eps, sign, use_ncp)   (dynamics.py:1398-1467)"""

    @staticmethod
    def forward(ctx, x, v, s, t, q, mask, eps, sign, use_ncp):
        ctx.save_for_backward(x, v, s, t, q, mask, eps)
        ctx.sign, ctx.use_ncp = sign, use_ncp
        out, logdet = ops.u1_xupdate(x.detach(), v.detach(), _opt(s), _opt(t), _opt(q), mask, float(eps), sign, use_ncp)
        return out, logdet

    @staticmethod
    def backward(ctx, gout, glogdet):
        x, v, s, t, q, mask, eps = ctx.saved_tensors
        gx, gv, gs, gt, gq, geps = ops.u1_xupdate_bwd(x.detach(), v.detach(), _opt(s), _opt(t), _opt(q), mask, float(eps),
                                                      ctx.sign, ctx.use_ncp, gout, glogdet)
        rs = lambda g, ref: None if (g is None or ref is None) else g.reshape(ref.shape).to(ref.dtype)  # noqa: E731
        return (gx.reshape(x.shape), gv.reshape(v.shape).to(v.dtype), rs(gs, s), rs(gt, t), rs(gq, q), None,
                _eps_grad(geps, eps), None, None)


# ---------------------------------------------------------------------------
# SU(3).  Gradients of complex tensors follow torch's convention
# (dL/dRe + i dL/dIm).  Like the reference, the force depends on x only through
# its explicit `@ x^+` factor (dsdx itself is not part of the graph: no
# create_graph, lattice/su3/pytorch/lattice.py:306).
#
# Two adjoints are not hand-written yet and are obtained by re-evaluating a
# differentiable torch restatement INSIDE backward (GPU ATen ops, no CPU path):
# projectSU (closed-form eigen/acos formula) and the per-site Wilson loops.
# ---------------------------------------------------------------------------
def _project_su_torch(x: Tensor) -> Tensor:
    """differentiable restatement of projectSU (group/su3/pytorch/utils.py:227-346)
    used only to back-propagate through l2b_su3_project"""
    eye = torch.eye(3, dtype=x.dtype, device=x.device)
    t = x.mH @ x
    t2 = t @ t
    tr = torch.diagonal(t, dim1=-2, dim2=-1).sum(-1).real
    p2 = torch.diagonal(t2, dim1=-2, dim2=-1).sum(-1).real
    det = torch.linalg.det(t).real
    tr3 = tr / 3.0
    tr32 = tr3 * tr3
    q = (0.5 * (p2 / 3.0 - tr32)).abs()
    r = 0.25 * tr3 * (5.0 * tr32 - p2) - 0.5 * det
    sq = q.sqrt()
    isq3 = (1.0 / (q * sq)).clamp(-3e38, 3e38)
    rsq3 = (r * isq3).clamp(-1.0, 1.0).clamp(-1.0 + 1e-12, 1.0 - 1e-12)
    th = torch.acos(rsq3) / 3.0
    sqc = sq * th.cos()
    sqs = (3.0 ** 0.5) * sq * th.sin()
    ll = tr3 + sqc
    e0, e1, e2 = tr3 - 2.0 * sqc, ll + sqs, ll - sqs
    s0, s1, s2 = e0.abs().sqrt(), e1.abs().sqrt(), e2.abs().sqrt()
    u = s0 + s1 + s2
    w = s0 * s1 * s2
    di = 1.0 / (w * (s0 + s1) * (s0 + s2) * (s1 + s2))
    c0 = di * (w * u * u + e0 * s0 * (e1 + e2) + e1 * s1 * (e0 + e2) + e2 * s2 * (e0 + e1))
    c1 = -(tr * u + w) * di
    c2 = u * di
    rs = c0[..., None, None] * eye + c1[..., None, None] * t + c2[..., None, None] * t2
    m = x @ rs
    d = torch.linalg.det(m)
    ph = -torch.atan2(d.imag, d.real) / 3.0
    return m * torch.complex(ph.cos(), ph.sin())[..., None, None]


def _wilson_loops_torch(x: Tensor) -> Tensor:
    """differentiable restatement of LatticeSU3._wilson_loops (lattice.py:157-199)"""
    ps = []
    for u in range(1, 4):
        for v in range(0, u):
            xu, xv = x[:, u], x[:, v]
            yuv = xu @ xv.roll(-1, dims=u + 1)
            yvu = xv @ xu.roll(-1, dims=v + 1)
            ps.append(torch.diagonal(yuv @ yvu.adjoint(), dim1=-2, dim2=-1).sum(-1))
    return torch.stack(ps)


def _vjp(fn, x: Tensor, g: Tensor) -> Tensor:
    with torch.enable_grad():
        xr = x.detach().requires_grad_(True)
        y = fn(xr)
        gx, = torch.autograd.grad(y, xr, grad_outputs=g.to(y.dtype))
    return gx


class SU3Project(torch.autograd.Function):
    """projectSU (compat_proj)"""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.su3_project(x.detach())

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        return _vjp(_project_su_torch, x, g)


class SU3GroupToVec(torch.autograd.Function):
    """su3_to_vec(projectSU(x)) in one kernel (vnet input packing, dynamics.py:1154-1156)"""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.su3_project(x.detach(), want_matrix=False, want_vec=True)

    @staticmethod
    def backward(ctx, gvec):
        x, = ctx.saved_tensors
        gy = ops.su3_to_vec_bwd(gvec)                 # adjoint of the linear vec8 map (kernel)
        return _vjp(_project_su_torch, x, gy)         # adjoint of projectSU (torch restatement)


class SU3WilsonLoops(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.su3_wilson_loops(x.detach())

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        return _vjp(_wilson_loops_torch, x, g)


class SU3Action(torch.autograd.Function):
    """S[b] = -(beta/3) sum Re tr P;  dS/dU = -(beta/3) A^+ (staple-sum kernel)"""

    @staticmethod
    def forward(ctx, x, beta):
        ctx.save_for_backward(x)
        ctx.beta = beta
        return ops.su3_plaq_sums(x.detach())[:, 0] * (-beta / 3.0)

    @staticmethod
    def backward(ctx, gs):
        x, = ctx.saved_tensors
        return ops.su3_action_grad(x.detach(), gs.to(torch.float64) * (-ctx.beta / 3.0)).reshape(x.shape), None


class SU3Force(torch.autograd.Function):
    """F = (beta/3) TAH(U A).  The reference builds it as projectTAH(dsdx @ x^+) with
    dsdx = autograd(S) NOT part of the graph but the explicit `@ x.adjoint()` still
    attached (lattice/su3/pytorch/lattice.py:299-308), so its backward is
    the partial derivative at fixed dsdx = -(beta/3) A^+:
        G_x = (TAH(G_F))^+ dsdx."""

    @staticmethod
    def forward(ctx, x, beta):
        ctx.save_for_backward(x)
        ctx.beta = beta
        return ops.su3_force(x.detach(), beta)

    @staticmethod
    def backward(ctx, gf):
        x, = ctx.saved_tensors
        gy = ops.su3_tah(gf.contiguous())                       # TAH is self-adjoint
        ones = torch.ones(x.shape[0], dtype=torch.float64, device=x.device)
        dsdx = ops.su3_action_grad(x.detach(), ones * (-ctx.beta / 3.0))
        return (gy.mH @ dsdx).reshape(x.shape), None


class SU3Kinetic(torch.autograd.Function):
    @staticmethod
    def forward(ctx, p):
        ctx.save_for_backward(p)
        return ops.su3_kinetic(p.detach())

    @staticmethod
    def backward(ctx, g):
        p, = ctx.saved_tensors
        pr = torch.view_as_real(p.detach().contiguous())
        return torch.view_as_complex(ops.rowscale(pr, g.to(torch.float64)).reshape(pr.shape)).reshape(p.shape)


class SU3VUpdate(torch.autograd.Function):
    """(v', logdet) = vupdate(v, F, s, t, q; eps, sign) on complex v   (dynamics.py:1266-1297)"""

    @staticmethod
    def forward(ctx, v, force, s, t, q, eps, sign):
        ctx.save_for_backward(v, force, s, t, q, eps)
        ctx.sign = sign
        return ops.su3_vupdate(v.detach(), force.detach(), _opt(s), _opt(t), _opt(q), float(eps), sign)

    @staticmethod
    def backward(ctx, gout, glogdet):
        v, force, s, t, q, eps = ctx.saved_tensors
        gv, gf, gs, gt, gq, geps = ops.su3_vupdate_bwd(v.detach(), force.detach(), _opt(s), _opt(t), _opt(q),
                                                       float(eps), ctx.sign, gout, glogdet)
        rs = lambda g, ref: None if (g is None or ref is None) else g.reshape(ref.shape).to(ref.dtype)  # noqa: E731
        return (gv.reshape(v.shape), gf.reshape(force.shape), rs(gs, s), rs(gt, t), rs(gq, q), _eps_grad(geps, eps),
                None)


class SU3UpdateGauge(torch.autograd.Function):
    """x' = m*x + exp(sign eps p) ((1-m)*x)   (dynamics.py:1420-1425,1468-1474)"""

    @staticmethod
    def forward(ctx, x, p, eps, mask, sign):
        ctx.save_for_backward(x, p, eps, mask)
        ctx.sign = sign
        return ops.su3_update_gauge(x.detach(), p.detach(), sign * float(eps), mask=mask)

    @staticmethod
    def backward(ctx, g):
        x, p, eps, mask = ctx.saved_tensors
        gx, gp, geps, bad = ops.su3_update_gauge_bwd(x.detach(), p.detach(), ctx.sign * float(eps), mask, False, g)
        if int(bad) != 0:
            raise ops.L2BError('matrix-exponential adjoint: ||eps p||_F > 3 (outside the series\' validated range)')
        return gx.reshape(x.shape), gp.reshape(p.shape), _eps_grad(geps, eps) * ctx.sign, None, None

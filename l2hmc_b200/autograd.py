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

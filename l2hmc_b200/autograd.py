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


_BAD_FLAGS: list = []


def check_exp_adjoint_flags() -> None:
    """raise if any matrix-exponential adjoint of the last backward pass saw ||eps p||_F > 3
    (one device read for the whole pass instead of a sync per x-update)"""
    if not _BAD_FLAGS:
        return
    bad = int(torch.stack([b.reshape(()) for b in _BAD_FLAGS]).sum())
    _BAD_FLAGS.clear()
    if bad != 0:
        raise ops.L2BError('matrix-exponential adjoint: ||eps p||_F > 3 (outside the series\' validated range)')


def _eps_grad(geps_per_chain: Tensor, eps: Tensor):
    return geps_per_chain.sum().to(eps.dtype).reshape(eps.shape)


class EpsPrime(torch.autograd.Function):
    """eps' = sigmoid(log eps) = eps / (1 + eps) (reference dynamics.py:1270,1394) of ALL step-size parameters of a
    sweep in one pass: 3 tiny launches forward and 4 backward per sweep instead of ~10 per update (the training step
    is launch-bound on the host; each update used to pay for its own log / sigmoid and their autograd nodes).
    Returns one 0-dim tensor per parameter, attached to the graph."""

    @staticmethod
    def forward(ctx, *params):
        p = torch.stack([q.detach().reshape(()) for q in params])
        e = torch.sigmoid(p.log())
        ctx.save_for_backward(p, e)
        ctx.shapes = [q.shape for q in params]
        # each step size in the parameters' dtype and as float64 (what the SU(3) kernels read)
        return tuple(e.unbind(0)) + tuple(e.to(torch.float64).unbind(0))

    @staticmethod
    def backward(ctx, *gs):
        p, e = ctx.saved_tensors
        n = p.shape[0]
        z = torch.zeros((), dtype=p.dtype, device=p.device)
        g = torch.stack([z if gi is None else gi.to(p.dtype).reshape(()) for gi in gs])
        gp = (g[:n] + g[n:]) * (e * (1.0 - e) / p)
        return tuple(gi.reshape(shp) for gi, shp in zip(gp.unbind(0), ctx.shapes))


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
    def forward(ctx, v, force, s, t, q, eps, sign, eps_value=None):
        ctx.save_for_backward(v, force, s, t, q, eps)
        ctx.sign = sign
        ctx.eps_value = eps.detach() if eps_value is None else eps_value   # 0-dim device tensor: read by the kernel
        out, logdet = ops.u1_vupdate(v.detach(), force.detach(), _opt(s), _opt(t), _opt(q), ctx.eps_value, sign)
        return out, logdet

    @staticmethod
    def backward(ctx, gout, glogdet):
        v, force, s, t, q, eps = ctx.saved_tensors
        gv, gf, gs, gt, gq, geps = ops.u1_vupdate_bwd(v.detach(), force.detach(), _opt(s), _opt(t), _opt(q),
                                                      ctx.eps_value, ctx.sign, gout, glogdet)
        rs = lambda g, ref: None if (g is None or ref is None) else g.reshape(ref.shape).to(ref.dtype)  # noqa: E731
        return (gv.reshape(v.shape), gf.reshape(force.shape).to(force.dtype), rs(gs, s), rs(gt, t), rs(gq, q),
                _eps_grad(geps, eps), None, None)


class U1XUpdate(torch.autograd.Function):
    """(x', logdet) = xupdate(x, v, s, t, q; mask, eps, sign, use_ncp)   (dynamics.py:1398-1467)"""

    @staticmethod
    def forward(ctx, x, v, s, t, q, mask, eps, sign, use_ncp, eps_value=None):
        ctx.save_for_backward(x, v, s, t, q, mask, eps)
        ctx.sign, ctx.use_ncp = sign, use_ncp
        ctx.eps_value = eps.detach() if eps_value is None else eps_value   # 0-dim device tensor: read by the kernel
        out, logdet = ops.u1_xupdate(x.detach(), v.detach(), _opt(s), _opt(t), _opt(q), mask, ctx.eps_value, sign,
                                     use_ncp)
        return out, logdet

    @staticmethod
    def backward(ctx, gout, glogdet):
        x, v, s, t, q, mask, eps = ctx.saved_tensors
        gx, gv, gs, gt, gq, geps = ops.u1_xupdate_bwd(x.detach(), v.detach(), _opt(s), _opt(t), _opt(q), mask,
                                                      ctx.eps_value, ctx.sign, ctx.use_ncp, gout, glogdet)
        rs = lambda g, ref: None if (g is None or ref is None) else g.reshape(ref.shape).to(ref.dtype)  # noqa: E731
        return (gx.reshape(x.shape), gv.reshape(v.shape).to(v.dtype), rs(gs, s), rs(gt, t), rs(gq, q), None,
                _eps_grad(geps, eps), None, None, None)


# ---------------------------------------------------------------------------
# SU(3).  Gradients of complex tensors follow torch's convention
# (dL/dRe + i dL/dIm).  Like the reference, the force depends on x only through
# its explicit `@ x^+` factor (dsdx itself is not part of the graph: no
# create_graph, lattice/su3/pytorch/lattice.py:306).
#
# Every adjoint is a hand-written kernel (include/l2b.h, "adjoints").
# ---------------------------------------------------------------------------
class SU3Project(torch.autograd.Function):
    """projectSU (compat_proj)"""

    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.su3_project(x.detach())

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        return ops.su3_project_bwd(x.detach(), gmat=g).reshape(x.shape)


class SU3GroupToVec(torch.autograd.Function):
    """su3_to_vec(projectSU(x)) in one kernel (vnet input packing, dynamics.py:1154-1156),
    written in `dtype` (the nets' element type) so no cast pass precedes the input GEMM;
    backward is the closed-form adjoint kernel (l2b_su3_project_bwd)"""

    @staticmethod
    def forward(ctx, x, dtype=torch.float64):
        ctx.save_for_backward(x)
        return ops.su3_project_vec(x.detach(), dtype)

    @staticmethod
    def backward(ctx, gvec):
        x, = ctx.saved_tensors
        return ops.su3_project_bwd(x.detach(), gvec=gvec).reshape(x.shape), None


class SU3WilsonLoops(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.save_for_backward(x)
        return ops.su3_wilson_loops(x.detach())

    @staticmethod
    def backward(ctx, g):
        x, = ctx.saved_tensors
        return ops.su3_wilson_loops_bwd(x.detach(), g).reshape(x.shape)


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
        x, = ctx.saved_tensors                                  # TAH is self-adjoint
        return ops.su3_force_bwd(x.detach(), ctx.beta, gf.contiguous()).reshape(x.shape), None


class SU3ActionC1(torch.autograd.Function):
    """improved action S[b] = -(beta/3) [(1 - 8 c1) sum Re tr P + c1 sum Re tr R]
    (lattice/su3/pytorch/lattice.py:96-112,252-269); dS/dU = -(beta/3) [(1 - 8 c1) A + c1 R]^+"""

    @staticmethod
    def forward(ctx, x, beta, c1):
        ctx.save_for_backward(x)
        ctx.beta, ctx.c1 = beta, c1
        sums = ops.su3_force_c1(x.detach(), beta, c1, want_force=False, want_sums=True)
        return ((1.0 - 8.0 * c1) * sums[:, 0] + c1 * sums[:, 1]) * (-beta / 3.0)

    @staticmethod
    def backward(ctx, gs):
        x, = ctx.saved_tensors
        gx = ops.su3_action_grad_c1(x.detach(), ctx.c1, coef=gs.to(torch.float64) * (-ctx.beta / 3.0))
        return gx.reshape(x.shape), None, None


class SU3ForceC1(torch.autograd.Function):
    """improved-action force, same graph semantics as SU3Force (dsdx detached, `@ x^+` attached)"""

    @staticmethod
    def forward(ctx, x, beta, c1):
        ctx.save_for_backward(x)
        ctx.beta, ctx.c1 = beta, c1
        return ops.su3_force_c1(x.detach(), beta, c1)

    @staticmethod
    def backward(ctx, gf):
        x, = ctx.saved_tensors
        gx = ops.su3_action_grad_c1(x.detach(), ctx.c1, scale=-ctx.beta / 3.0, gforce=gf.contiguous())
        return gx.reshape(x.shape), None, None


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
    def forward(ctx, v, force, s, t, q, eps, sign, eps_value=None):
        ctx.save_for_backward(v, force, s, t, q, eps)
        ctx.sign = sign
        ctx.eps_value = eps.detach() if eps_value is None else eps_value   # 0-dim device tensor: read by the kernel
        return ops.su3_vupdate(v.detach(), force.detach(), _opt(s), _opt(t), _opt(q), ctx.eps_value, sign)

    @staticmethod
    def backward(ctx, gout, glogdet):
        v, force, s, t, q, eps = ctx.saved_tensors
        gv, gf, gs, gt, gq, geps = ops.su3_vupdate_bwd(v.detach(), force.detach(), _opt(s), _opt(t), _opt(q),
                                                       ctx.eps_value, ctx.sign, gout, glogdet)
        rs = lambda g, ref: None if (g is None or ref is None) else g.reshape(ref.shape).to(ref.dtype)  # noqa: E731
        return (gv.reshape(v.shape), gf.reshape(force.shape), rs(gs, s), rs(gt, t), rs(gq, q), _eps_grad(geps, eps),
                None, None)


# While True (Trainer.train_step sets it around loss.backward()), the weight / bias gradients of the
# vnet heads are not formed in every SU3HeadsVUpdate.backward: the pre-activation cotangents and z of all
# the v-updates that share a network are stashed and ONE dW = cat(G)^T cat(z) per head is formed when the
# backward pass ends (autograd engine callback) -- instead of 2*N_LF*2 GEMMs each followed by a bf16->fp32
# cast and an accumulation into a 75 MB .grad (8^4: ~38 GB of traffic per training step saved).
DEFER_HEAD_GRADS = False
# Multi-rank training (Trainer.train_step sets it around loss.backward()): a dist.GradBucket.  The deferred dW GEMMs
# then write straight into the bucket's flat exchange buffer and start the all-reduce of their slice at once, so the
# collective of one head overlaps the GEMM of the next and the rest of the backward pass.
HEAD_GRAD_SINK = None


def _weight_grad_segments(param, gs, xs, sink, mode: str = 'auto') -> None:
    """dW = sum_u g_u^T x_u for ONE weight matrix from the stashed (cotangent, input) pairs of all the updates that
    share it: one tensor-core launch (l2b_gemm_bf16, up to 32 segments; both operands contracted over their rows, no
    concatenation and no transposed copy) written straight into the gradient exchange buffer when there is one
    (its all-reduce starts at once), else into / onto `param.grad`.  mode 'x3': the operands are bf16x3 splits of
    fp32 matrices (fp32-accurate product, ops.gemm_f32).  Non-bf16 cotangents in mode 'auto' (fp64 nets) keep the
    library GEMM."""
    if not param.requires_grad:
        return
    x3 = mode == 'x3'
    bf16 = (not x3) and all(t.dtype == torch.bfloat16 for t in gs + xs)
    use_sink = sink is not None and param.grad is None and id(param) in sink.offsets

    def form(out, out_dtype):
        if x3:
            res = ops.gemm_f32(gs, xs, False, False)
            return res[:param.shape[0], :param.shape[1]]
        if bf16:
            for i in range(0, len(gs), 32):
                out = ops.gemm_bf16(gs[i:i + 32], xs[i:i + 32], False, False, out=out, out_dtype=out_dtype,
                                    accumulate=i > 0)
            return out
        g, x = torch.cat(gs), torch.cat(xs)
        res = (g.t() @ x).to(out_dtype)
        if out is not None:
            out.copy_(res)
            return out
        return res
    if use_sink:
        out = sink.view(param)
        if bf16 and out.dtype in (torch.bfloat16, torch.float32):
            form(out, out.dtype)
        else:
            out.copy_(form(None, torch.float32))
        sink.reduce_async(param)             # its all-reduce starts while the next weight's GEMM runs
        return
    g = form(None, torch.float32).to(param.dtype).reshape(param.shape)
    param.grad = g if param.grad is None else param.grad + g


def _flush_head_grads(net) -> None:
    pend = net._pending_head_grads
    net._pending_head_grads = None
    if not pend:
        return
    ws, bs, cs, wt, bt, wq, bq, cq = net.head_params()
    zs = [p[3] for p in pend]
    sink = HEAD_GRAD_SINK

    def acc(param, g):
        if param.requires_grad:
            g = g.to(param.dtype).reshape(param.shape)
            param.grad = g if param.grad is None else param.grad + g
    for k, (w, b_) in enumerate(((ws, bs), (wt, bt), (wq, bq))):
        gs = [p[k] for p in pend]                        # per update: [nb, xdim] pre-activation cotangents
        _weight_grad_segments(w, gs, zs, sink)
    # bias and ScaledTanh.coeff gradients: the adjoint kernel's own sums over the chains, summed over the updates
    col = torch.stack([p[4] for p in pend]).sum(0)       # [5, xdim]
    for k, b_ in enumerate((bs, bt, bq)):
        acc(b_, col[k])
    acc(cs, col[3:4])
    acc(cq, col[4:5])


# ---------------------------------------------------------------------------
# dense layers of the networks on the tensor-core GEMM (csrc/l2b_gemm.cu)
# ---------------------------------------------------------------------------
_PENDING_DENSE: dict = {}         # id(weight) -> (weight, [cotangents], [inputs]) of the current backward pass


def _flush_dense_grads() -> None:
    pend = dict(_PENDING_DENSE)
    _PENDING_DENSE.clear()
    for w, gs, xs, mode in pend.values():
        _weight_grad_segments(w, gs, xs, HEAD_GRAD_SINK, mode)


def _act_grad(act, y: Tensor, pre, g: Tensor) -> Tensor:
    """cotangent of the pre-activation from the cotangent of y = act(pre); tanh / relu / leaky_relu / elu from the
    output alone (what ATen's own backward formulas use), swish from the stored pre-activation"""
    g = g.float()
    if act in (None, 'identity'):
        return g
    yf = y.float()
    if act == 'tanh':
        return g * (1.0 - yf * yf)
    if act == 'relu':
        return g * (yf > 0)
    if act == 'leaky_relu':
        return g * torch.where(yf > 0, 1.0, 0.01)
    if act == 'elu':
        return g * torch.where(yf > 0, 1.0, yf + 1.0)
    if act == 'swish':
        pf = pre.float()
        sg = torch.sigmoid(pf)
        return g * (sg * (1.0 + pf * (1.0 - sg)))
    raise ValueError(f'unknown activation {act!r}')


class TCDense(torch.autograd.Function):
    """y = act(sum_i x_i W_i^T + sum_i b_i): ONE Linear (a hidden layer, reference network/pytorch/network.py:538-541;
    an output head :546-548) or the InputLayer's pair sharing one output (:415-451), as a single launch of the
    tensor-core GEMM with bias and activation in its epilogue.  mode 'bf16': bf16 operands, fp32 accumulation (what
    autocast runs); mode 'x3': fp32 nets without autocast -- every operand as its bf16x3 split, six products per pair,
    fp32-accurate.  Backward: dX_i = (g act') W_i and dW_i = (g act')^T x_i on the same kernel (W_i and the cotangent
    contracted over their rows: no transposed copies); under DEFER_HEAD_GRADS the dW of all the updates sharing a
    weight matrix are formed by one launch when the backward pass ends.  args = (x_0, W_0, b_0[, x_1, W_1, b_1])."""

    @staticmethod
    def forward(ctx, act, owner, mode, *args):
        n = len(args) // 3
        xs, ws, bs = [args[3 * i] for i in range(n)], [args[3 * i + 1] for i in range(n)], [args[3 * i + 2] for i in range(n)]
        bias = None
        for b_ in bs:
            if b_ is not None:
                bias = b_.detach().float() if bias is None else bias + b_.detach().float()
        fused = None if act == 'swish' else act
        nout = int(ws[0].shape[0])
        # inputs of different widths (the U(1) xnet: 4TX cos / sin values against 2TX momenta) cannot be segments of
        # one launch: they are concatenated along K instead (small matrices; the SU(3) pair shares its shape)
        ragged = n == 2 and xs[0].numel() // xs[0].shape[0] != xs[1].numel() // xs[1].shape[0]
        if mode == 'x3':
            xb = [ops.split_bf16x3(x.detach().reshape(x.shape[0], -1).float()) for x in xs]
            wb = [owner.weight_split3(w) for w in ws]
            if ragged:
                y = ops.gemm_f32(torch.cat(xb, 2), torch.cat(wb, 2), True, True, bias=bias, act=fused)
            else:
                y = ops.gemm_f32(xb, wb, True, True, bias=bias, act=fused)
            if y.shape[1] != nout:
                y = y[:, :nout].contiguous()
        else:
            xb = [x.detach().reshape(x.shape[0], -1).to(torch.bfloat16) for x in xs]
            wb = [owner.weight_as_bf16(w) for w in ws]
            if ragged:
                y = ops.gemm_bf16(torch.cat(xb, 1), torch.cat(wb, 1), True, True, bias=bias, act=fused)
            else:
                y = ops.gemm_bf16(xb, wb, True, True, bias=bias, act=fused)
        pre = None
        if act == 'swish':
            pre, y = y, torch.nn.functional.silu(y)
        ctx.act, ctx.n, ctx.mode = act, n, mode
        ctx.meta = [(x.shape, x.dtype) for x in xs]
        ctx.save_for_backward(y, pre, *xb, *wb, *ws, *[b_ for b_ in bs])
        return y

    @staticmethod
    def backward(ctx, gy):
        n, x3 = ctx.n, ctx.mode == 'x3'
        y, pre = ctx.saved_tensors[:2]
        xb = ctx.saved_tensors[2:2 + n]
        wb = ctx.saved_tensors[2 + n:2 + 2 * n]
        ws = ctx.saved_tensors[2 + 2 * n:2 + 3 * n]
        bs = ctx.saved_tensors[2 + 3 * n:2 + 4 * n]
        gpre32 = _act_grad(ctx.act, y, pre, gy)
        gpre = ops.split_bf16x3(gpre32.contiguous()) if x3 else gpre32.to(torch.bfloat16)
        grads = []
        gb = None
        for i in range(n):
            need_x, need_w, need_b = ctx.needs_input_grad[3 + 3 * i:6 + 3 * i]
            gx = gw = gbi = None
            if need_x:
                shp, dt = ctx.meta[i]
                if x3:
                    gx = ops.gemm_f32(gpre, wb[i], True, False)
                    nin = int(ws[i].shape[1])
                    gx = (gx if gx.shape[1] == nin else gx[:, :nin]).reshape(shp).to(dt)
                else:
                    gx = ops.linear_dx(gpre, wb[i]).reshape(shp).to(dt)
            if need_w:
                if DEFER_HEAD_GRADS:
                    if not _PENDING_DENSE:
                        torch.autograd.Variable._execution_engine.queue_callback(_flush_dense_grads)
                    ent = _PENDING_DENSE.setdefault(id(ws[i]), (ws[i], [], [], ctx.mode))
                    ent[1].append(gpre)
                    ent[2].append(xb[i])
                elif x3:
                    gw = ops.gemm_f32(gpre, xb[i], False, False)[:ws[i].shape[0], :ws[i].shape[1]].to(ws[i].dtype)
                else:
                    gw = ops.linear_dw(gpre, xb[i]).to(ws[i].dtype)
            if need_b and bs[i] is not None:
                if gb is None:
                    gb = gpre32.sum(0)
                gbi = gb.to(bs[i].dtype).reshape(bs[i].shape)
            grads += [gx, gw, gbi]
        return (None, None, None, *grads)


class ConvPeriodic(torch.autograd.Function):
    """One block of the U(1) xnet's ConvStack (reference network/pytorch/network.py:296-313): PeriodicPadding(n - 1)
    + Conv2d(f, n) [+ the activation, when no pooling sits in between] as gather -> tensor-core GEMM
    (ops.conv_im2col, ops.gemm_bf16 / gemm_f32); output NHWC.  Backward: the gathered matrix is rebuilt from the saved
    input (it is 25 - 100 x the input's size), dW = g^T col (split-K over the nb OH OW rows), dcol = g W, and the
    gather's adjoint (ops.conv_col2im).  On NHWC inputs with Cin % 8 == 0 (every block after the first of the default
    stack) the gathered columns are tap-major, k = (kh, kw, ci): the gather moves 32-byte channel vectors and the
    weight is viewed as [Cout, n, n, Cin] (cached image); dW comes back in that order and is permuted once."""

    @staticmethod
    def forward(ctx, x, weight, bias, nchw, mode, act, owner):
        n = int(weight.shape[-1])
        cout = int(weight.shape[0])
        x3 = mode in ('x3', 'x2')
        planes = {'x3': 3, 'x2': 2}.get(mode, 1)
        xd = x.detach()
        if xd.dtype not in (torch.float32, torch.bfloat16):
            xd = xd.float()
        if not x3 and xd.dtype != torch.bfloat16:
            xd = xd.to(torch.bfloat16)
        tap = (not nchw) and int(weight.shape[1]) % 8 == 0
        col = ops.conv_im2col(xd, n, nchw, planes, tap)
        fused = None if act == 'swish' else act
        b_ = None if bias is None else bias.detach().float()
        if x3:
            y = ops.gemm_f32(col, owner.weight_split3(weight, tap)[:planes], True, True, bias=b_, act=fused)
        else:
            y = ops.gemm_bf16(col[0], owner.weight_as_bf16(weight, tap), True, True, bias=b_, act=fused)
        if y.shape[1] != cout:
            y = y[:, :cout].contiguous()
        nb = int(x.shape[0])
        H, W = (int(x.shape[2]), int(x.shape[3])) if nchw else (int(x.shape[1]), int(x.shape[2]))
        y = y.reshape(nb, H + n - 1, W + n - 1, cout)
        pre = None
        if act == 'swish':
            pre, y = y, torch.nn.functional.silu(y)
        ctx.cfg = (n, nchw, mode, act, owner, tap)
        ctx.xmeta = (x.shape, x.dtype)
        ctx.save_for_backward(xd, weight, bias, y if act is not None else None, pre)
        return y

    @staticmethod
    def backward(ctx, gy):
        n, nchw, mode, act, owner, tap = ctx.cfg
        xd, weight, bias, y, pre = ctx.saved_tensors
        x3 = mode in ('x3', 'x2')
        planes = {'x3': 3, 'x2': 2}.get(mode, 1)
        cout = int(weight.shape[0])
        K = int(weight.numel() // cout)
        g32 = _act_grad(act, y, pre, gy).reshape(-1, cout).contiguous()
        gx = gw = gb = None
        need_x, need_w, need_b = ctx.needs_input_grad[0], ctx.needs_input_grad[1], ctx.needs_input_grad[2]
        g = ops.split_bf16x3(g32)[:planes] if x3 else g32.to(torch.bfloat16)
        if need_w:
            col = ops.conv_im2col(xd, n, nchw, planes, tap)
            if x3:
                gw = ops.gemm_f32(g, col, False, False)
            else:
                gw = ops.gemm_bf16(g, col[0], False, False, out_dtype=torch.float32)
            if tap:
                gw = gw[:cout, :K].reshape(cout, n, n, K // (n * n)).permute(0, 3, 1, 2).to(weight.dtype).contiguous()
            else:
                gw = gw[:cout, :K].reshape(weight.shape).to(weight.dtype)
            del col
        if need_x:
            if x3:
                dcol = ops.gemm_f32(g, owner.weight_split3(weight, tap)[:planes], True, False)
            else:
                dcol = ops.gemm_bf16(g, owner.weight_as_bf16(weight, tap), True, False, out_dtype=torch.bfloat16)
            gx = ops.conv_col2im(dcol, xd, n, nchw, tap).to(ctx.xmeta[1]).reshape(ctx.xmeta[0])
        if need_b and bias is not None:
            gb = g32.sum(0).to(bias.dtype)
        return gx, gw, gb, None, None, None, None


class PoolAct(torch.autograd.Function):
    """MaxPool2d(p) followed by the activation (network.py:314-322) on NHWC, one kernel each way"""

    @staticmethod
    def forward(ctx, x, pool, act):
        y, idx, pre = ops.pool_act(x.detach(), pool, act)
        ctx.cfg = (pool, act, x.shape, x.dtype)
        ctx.save_for_backward(y, idx, pre)
        return y

    @staticmethod
    def backward(ctx, gy):
        pool, act, shape, dt = ctx.cfg
        y, idx, pre = ctx.saved_tensors
        return ops.pool_act_bwd(gy, y, pre, idx, shape, pool, act).to(dt), None, None


class SU3HeadsVUpdate(torch.autograd.Function):
    """(v', logdet) = vupdate(v, F, heads(z); eps, sign) with the three head GEMMs on the
    tensor cores and s, t, q kept on chip (l2b_su3_heads_vupdate, csrc/l2b_vnet.cu).
    When a gradient is needed the kernel also writes (s, t, q) once (fp32) for the backward pass:
    ONE element-wise adjoint kernel (l2b_su3_heads_vupdate_bwd) then yields gv, gF, d/d eps and the
    cotangents of the heads' pre-activations in the GEMM dtype, and the GEMMs of the Linear backward run on
    the tensor-core GEMM l2b_gemm_bf16 (dz now: one split-K launch over the three heads; dW either now or, under
    DEFER_HEAD_GRADS, once per backward pass for all the v-updates sharing the net); fp32 / fp64 nets without
    autocast keep the library GEMM."""

    @staticmethod
    def forward(ctx, z, v, force, eps, sign, eps_value, net, *head_params):
        ctx.sign, ctx.net = sign, net
        ctx.eps_value = eps.detach() if eps_value is None else eps_value   # 0-dim device tensor: read by the kernel
        ctx.autocast = (torch.is_autocast_enabled('cuda'), torch.get_autocast_dtype('cuda'))
        need = any(ctx.needs_input_grad)
        pack = net.heads_pack()
        res = ops.su3_heads_vupdate(z.detach(), pack, v.detach(), force.detach(), ctx.eps_value, sign, want_stq=need)
        ctx.pack = pack
        ctx.save_for_backward(z, v, force, eps, res[2] if need else None, *head_params)
        return res[0], res[1]

    @staticmethod
    def backward(ctx, gout, glogdet):
        z, v, force, eps, stq = ctx.saved_tensors[:5]
        ws, bs, cs, wt, bt, wq, bq, cq = ctx.saved_tensors[5:]
        net = ctx.net
        cdt = ctx.autocast[1] if ctx.autocast[0] else ws.dtype        # the dtype the reference's Linear backward runs in
        if cdt not in (torch.float32, torch.bfloat16):
            cdt = torch.float32
        # one element-wise kernel: gv, gF, d/d eps, the pre-activation cotangents of the three heads in the
        # GEMM dtype, and gs*s / gq*q for the ScaledTanh.coeff gradients
        gv, gf, gpre, colsum, geps = ops.su3_heads_vupdate_bwd(v.detach(), force.detach(), stq, ctx.pack,
                                                               ctx.eps_value, ctx.sign, gout, glogdet, cdt,
                                                               want_gforce=ctx.needs_input_grad[2])
        gp = (gpre[0], gpre[1], gpre[2])
        wc = net.head_weights_as(cdt)
        if cdt == torch.bfloat16:
            # dz = sum_heads g_h W_h: one launch, three segments, the weights contracted over their rows (split-K)
            gz = ops.linear_dx(list(gp), list(wc)).to(z.dtype)
        else:
            gz = (gp[0] @ wc[0] + gp[1] @ wc[1] + gp[2] @ wc[2]).to(z.dtype)
        zc = z.detach().to(cdt)
        if DEFER_HEAD_GRADS:
            if getattr(net, '_pending_head_grads', None) is None:
                net._pending_head_grads = []
                torch.autograd.Variable._execution_engine.queue_callback(lambda: _flush_head_grads(net))
            net._pending_head_grads.append((gp[0], gp[1], gp[2], zc, colsum))
            gparams = (None,) * 8
        else:
            need = ctx.needs_input_grad[7:]
            if cdt == torch.bfloat16:
                gw = [ops.linear_dw(g, zc).to(w.dtype) if nd else None
                      for g, w, nd in zip(gp, (ws, wt, wq), (need[0], need[3], need[5]))]
            else:
                gw = [(g.t() @ zc).to(w.dtype) if nd else None
                      for g, w, nd in zip(gp, (ws, wt, wq), (need[0], need[3], need[5]))]
            gb = [colsum[k].to(b_.dtype) if nd else None
                  for k, (b_, nd) in enumerate(zip((bs, bt, bq), (need[1], need[4], need[6])))]
            gparams = (gw[0], gb[0], colsum[3:4].to(cs.dtype) if need[2] else None, gw[1], gb[1], gw[2], gb[2],
                       colsum[4:5].to(cq.dtype) if need[7] else None)
        return (gz, gv.reshape(v.shape), None if gf is None else gf.reshape(force.shape), _eps_grad(geps, eps), None,
                None, None, *gparams)


class SU3UpdateGauge(torch.autograd.Function):
    """x' = m*x + exp(sign eps p) ((1-m)*x)   (dynamics.py:1420-1425,1468-1474)"""

    @staticmethod
    def forward(ctx, x, p, eps, mask, sign, eps_value=None):
        ctx.save_for_backward(x, p, eps, mask)
        ctx.sign = sign
        ctx.eps_value = eps.detach() if eps_value is None else eps_value   # 0-dim device tensor: read by the kernel
        return ops.su3_update_gauge(x.detach(), p.detach(), ctx.eps_value, mask=mask, eps_mult=float(sign))

    @staticmethod
    def backward(ctx, g):
        x, p, eps, mask = ctx.saved_tensors
        gx, gp, geps, bad = ops.su3_update_gauge_bwd(x.detach(), p.detach(), ctx.eps_value, mask, False, g,
                                                     eps_mult=float(ctx.sign))
        _BAD_FLAGS.append(bad)        # checked once per backward pass (check_exp_adjoint_flags), not per link update
        if len(_BAD_FLAGS) > 1024:    # callers that never check: bound the list
            check_exp_adjoint_flags()
        return gx.reshape(x.shape), gp.reshape(p.shape), _eps_grad(geps, eps) * ctx.sign, None, None, None

"""TEST INFRASTRUCTURE (never the product path): stand-ins for the U(1) entry points of `l2hmc_b200.ops`, written
in CPU torch from the formulas of the oracle (oracle/u1.py, oracle/dynamics.py), plus the patches that let the
`Dynamics` mirror be constructed without a CUDA device.  With them the HOST logic of the mirror -- control flow
of the transition kernels, histories, masks, accept / reject mixing, the metrics contract -- runs on the CPU
against the goldens, so mistakes there are found before a GPU minute is spent.  The kernels themselves are
covered by the `-m gpu` tier; nothing outside tests/ imports this module."""
from __future__ import annotations

import contextlib
import math

import torch

PI, TWO_PI = math.pi, 2.0 * math.pi


def _field(x, shape=None):
    if x.dim() != 4:
        x = x.reshape(x.shape[0], 2, *shape)
    return x


def _e(eps, like):
    if isinstance(eps, torch.Tensor):
        return eps.to(like.dtype) if eps.requires_grad else eps.detach().to(like.dtype)
    return torch.tensor(float(eps), dtype=like.dtype)


def u1_wilson_loops(x, shape=None):
    x = _field(x, shape)
    return x[:, 0] + x[:, 1].roll(-1, 1) - x[:, 0].roll(-1, 2) - x[:, 1]


def u1_observables(x, beta, shape=None):
    w = u1_wilson_loops(x, shape)
    intq = (w - TWO_PI * torch.floor((w + PI) / TWO_PI)).sum((1, 2)) / TWO_PI
    return torch.stack([beta * (1.0 - w.cos()).sum((1, 2)), w.cos().mean((1, 2)), w.sin().sum((1, 2)) / TWO_PI, intq], 1)


def u1_force(x, beta, shape=None):
    x4 = _field(x, shape)
    s = u1_wilson_loops(x4).sin()
    return beta * torch.stack([s - s.roll(1, 2), -s + s.roll(1, 1)], 1)


def u1_kinetic(v):
    return 0.5 * (v.reshape(v.shape[0], -1) ** 2).sum(1)


def u1_compat_proj(x):
    return ((x + PI) % TWO_PI) - PI


def _rows(a, nb, like):
    return None if a is None else a.to(like.dtype).reshape(nb, -1)


def u1_vupdate(v, force, s, t, q, eps, sign):
    nb = v.shape[0]
    v2, f2 = v.reshape(nb, -1), force.to(v.dtype).reshape(nb, -1)
    e = _e(eps, v2)
    z = torch.zeros_like(v2)
    s, t, q = (z if a is None else _rows(a, nb, v2) for a in (s, t, q))
    kick = 0.5 * e * (f2 * torch.exp(e * q) + t)
    if sign > 0:
        out, logdet = torch.exp(0.5 * e * s) * v2 - kick, (0.5 * e * s).sum(1)
    else:
        out, logdet = torch.exp(-0.5 * e * s) * (v2 + kick), -(0.5 * e * s).sum(1)
    return out.reshape(v.shape), logdet


def u1_xupdate(x, v, s, t, q, mask, eps, sign, use_ncp):
    nb = x.shape[0]
    x2, v2 = x.reshape(nb, -1), v.to(x.dtype).reshape(nb, -1)
    e = _e(eps, x2)
    z = torch.zeros_like(x2)
    s, t, q = (z if a is None else _rows(a, nb, x2) for a in (s, t, q))
    m = mask.to(x2.dtype).reshape(1, -1)
    mb = 1.0 - m
    if sign > 0:
        if use_ncp:
            es = torch.exp(e * s)
            halfx = 0.5 * x2
            xp = 2.0 * torch.atan(torch.tan(halfx) * es)
            y = xp + e * (v2 * torch.exp(e * q) + t)
            logdet = (mb * torch.log(es / (halfx.cos() ** 2 + (es * halfx.sin()) ** 2))).sum(1)
        else:
            y = x2 * torch.exp(e * s) + e * (v2 * torch.exp(e * q) + t)
            logdet = (mb * e * s).sum(1)
    else:
        if use_ncp:                       # the reference's inverse, as written (dynamics.py:1452-1462)
            es = torch.exp(-e * s)
            halfx = 0.5 * x2
            y = 2.0 * torch.atan(es * torch.tan(halfx)) - es * e * (v2 * torch.exp(e * q) + t)
            logdet = (mb * torch.log(es / (halfx.cos() ** 2 + (es * halfx.sin()) ** 2))).sum(1)
        else:
            y = torch.exp(-e * s) * (x2 - e * (v2 * torch.exp(e * q) + t))
            logdet = -(mb * e * s).sum(1)
    out = u1_compat_proj(m * x2 + mb * y)
    return out.reshape(x.shape), logdet


def u1_hmc_trajectory(x, v, beta, eps, nlf, shape=None):
    x4 = _field(x, shape)
    nb = x4.shape[0]
    xs, vs = x4.reshape(nb, -1).clone(), v.reshape(nb, -1).clone()
    en = torch.empty(nb, 4, dtype=x4.dtype)
    en[:, 0], en[:, 1] = u1_kinetic(vs), u1_observables(x4, beta)[:, 0]
    for _ in range(nlf):
        vs = vs - 0.5 * eps * u1_force(xs.reshape(x4.shape), beta).reshape(nb, -1)
        xs = xs + eps * vs
        vs = vs - 0.5 * eps * u1_force(xs.reshape(x4.shape), beta).reshape(nb, -1)
    en[:, 2], en[:, 3] = u1_kinetic(vs), u1_observables(xs.reshape(x4.shape), beta)[:, 0]
    return xs.reshape(x4.shape), vs.reshape(x4.shape), en


def accept_mix(accept, pairs):
    nb = accept.numel()
    sel = accept.reshape(nb, 1) > 0
    return [torch.where(sel, b.to(a.dtype).reshape(nb, -1), a.reshape(nb, -1)) for a, b in pairs]


# ---- adjoints of the U(1) stand-ins: vector-Jacobian products by torch autograd of the functions above ----
def _leaf(a):
    return None if a is None else a.detach().clone().requires_grad_(True)


def _vjp(outs, gouts, leaves):
    pairs = [(o, g) for o, g in zip(outs, gouts) if g is not None]
    live = [a for a in leaves if a is not None]
    grads = iter(torch.autograd.grad([o for o, _ in pairs], live, grad_outputs=[g.to(o.dtype).reshape(o.shape) for o, g in pairs],
                                     allow_unused=True))
    res = []
    for a in leaves:
        if a is None:
            res.append(None)
        else:
            g = next(grads)
            res.append(torch.zeros_like(a) if g is None else g)
    return res


def u1_wilson_loops_bwd(gw):
    nb, T, X = gw.shape
    with torch.enable_grad():
        x = torch.zeros(nb, 2, T, X, dtype=gw.dtype, requires_grad=True)
        return _vjp([u1_wilson_loops(x)], [gw], [x])[0]


def u1_force_bwd(x, beta, gforce, shape=None):
    with torch.enable_grad():
        x4 = _leaf(_field(x, shape))
        return _vjp([u1_force(x4, beta)], [gforce.reshape(x4.shape)], [x4])[0]


def rowscale(a, scale):
    return a * scale.to(a.dtype).reshape(-1, *([1] * (a.dim() - 1)))


def u1_vupdate_bwd(v, force, s, t, q, eps, sign, gv_out, glogdet):
    nb = v.shape[0]
    with torch.enable_grad():
        lv, lf = _leaf(v.reshape(nb, -1)), _leaf(force.to(v.dtype).reshape(nb, -1))
        ls, lt, lq = (_leaf(_rows(a, nb, lv)) for a in (s, t, q))
        e = _e(eps, lv).expand(nb).clone().requires_grad_(True)          # one step size per chain -> per-chain d/d eps
        out, logdet = u1_vupdate(lv, lf, ls, lt, lq, e[:, None], sign)
        gv, gf, gs, gt, gq, ge = _vjp([out, logdet], [gv_out, glogdet], [lv, lf, ls, lt, lq, e])
    return gv, gf, gs, gt, gq, ge


def u1_xupdate_bwd(x, v, s, t, q, mask, eps, sign, use_ncp, gx_out, glogdet):
    nb = x.shape[0]
    with torch.enable_grad():
        lx, lv = _leaf(x.reshape(nb, -1)), _leaf(v.to(x.dtype).reshape(nb, -1))
        ls, lt, lq = (_leaf(_rows(a, nb, lx)) for a in (s, t, q))
        e = _e(eps, lx).expand(nb).clone().requires_grad_(True)
        out, logdet = u1_xupdate(lx, lv, ls, lt, lq, mask, e[:, None], sign, use_ncp)
        gx, gv, gs, gt, gq, ge = _vjp([out, logdet], [gx_out, glogdet], [lx, lv, ls, lt, lq, e])
    return gx, gv, gs, gt, gq, ge


class _FakeCuda:
    @staticmethod
    def is_available():
        return True

    @staticmethod
    def current_device():
        return 0


class _TorchProxy:
    """`torch` as seen by the dynamics module: everything real, except that 'the CUDA device' is the CPU"""
    cuda = _FakeCuda()

    @staticmethod
    def device(*args, **kwargs):
        return torch.device('cpu')

    def __getattr__(self, name):
        return getattr(torch, name)


# ---------------------------------------------------------------------------
# SU(3): thin torch <-> numpy wrappers around the oracle (oracle/su3.py, oracle/dynamics.py)
# ---------------------------------------------------------------------------
def _np(t):
    return t.detach().cpu().numpy()


def _full(x, dims=None):
    return x if x.dim() == 8 else x.reshape(x.shape[0], 4, *dims, 3, 3)


def su3_plaq_sums(x):
    from oracle import su3 as o
    re, im = o.plaq_sums(_np(x))
    return torch.stack([torch.from_numpy(re), torch.from_numpy(im)], 1)


def su3_wilson_loops(x):
    from oracle import su3 as o
    return torch.from_numpy(o.wilson_loops(_np(x)))


def su3_force(x, beta, want_plaq_sum=False):
    from oracle import su3 as o
    f = torch.from_numpy(o.grad_action(_np(x), float(beta)))
    return (f, su3_plaq_sums(x)[:, 0]) if want_plaq_sum else f


def su3_force_c1(x, beta, c1, want_force=True, want_sums=False):
    """improved action: the force through torch autograd of the torch restatement of the reference's loops"""
    from oracle import su3 as o
    xn = _np(x)
    sums = torch.stack([torch.from_numpy(o.plaq_sums(xn)[0]),
                        torch.from_numpy(o.rect_traces(xn).real.reshape(12, xn.shape[0], -1).sum(2).sum(0))], 1)
    if not want_force:
        return sums
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    lat = LatticeSU3(xn.shape[0], list(xn.shape[2:6]), c1=c1)
    lat.rect_kernel = False
    with torch.enable_grad():
        xr = x.detach().clone().requires_grad_(True)
        tr = lambda a: torch.diagonal(a, dim1=-2, dim2=-1).sum(-1)  # noqa: E731
        ps = 0.0
        for u in range(1, 4):
            for v in range(u):
                ps = ps + tr(lat._plaquette(xr, u, v)).real.flatten(1).sum(1)
        act = -(beta / 3.0) * (1 - 8 * c1) * ps + lat._rect_action(xr, beta)
        dsdx, = torch.autograd.grad(act.sum(), xr)
    y = dsdx @ x.mH
    a = 0.5 * (y - y.mH)
    f = a - torch.diagonal(a, dim1=-2, dim2=-1).sum(-1)[..., None, None] / 3.0 * torch.eye(3, dtype=a.dtype)
    return (f, sums) if want_sums else f


def _improved_staples_adjoint(x, c1):
    """Aimp^+ = [(1 - 8 c1) A + c1 R]^+ from torch autograd: dS/dx = -(beta/3) Aimp^+ at beta = 3 gives -Aimp^+"""
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    lat = LatticeSU3(x.shape[0], list(x.shape[2:6]), c1=c1)
    with torch.enable_grad():
        xr = x.detach().clone().requires_grad_(True)
        tr = lambda a: torch.diagonal(a, dim1=-2, dim2=-1).sum(-1)  # noqa: E731
        ps = 0.0
        for u in range(1, 4):
            for v in range(u):
                ps = ps + tr(lat._plaquette(xr, u, v)).real.flatten(1).sum(1)
        act = -(1 - 8 * c1) * ps + lat._rect_action(xr, 3.0)
        g, = torch.autograd.grad(act.sum(), xr)
    return -g


def su3_action_grad_c1(x, c1, coef=None, scale=0.0, gforce=None):
    ah = _improved_staples_adjoint(x, c1)
    if gforce is None:
        return coef.to(torch.float64).reshape(-1, *([1] * (x.dim() - 1))) * ah
    gf = gforce.reshape(x.shape)
    a = 0.5 * (gf - gf.mH)
    th = a - torch.diagonal(a, dim1=-2, dim2=-1).sum(-1)[..., None, None] / 3.0 * torch.eye(3, dtype=a.dtype)
    return th.mH @ (scale * ah)


def su3_action_grad(x, coef):
    return su3_action_grad_c1(_full(x), 0.0, coef=coef)


def su3_force_bwd(x, beta, gforce):
    return su3_action_grad_c1(_full(x), 0.0, scale=-float(beta) / 3.0, gforce=gforce)


def su3_project(x, want_matrix=True, want_vec=False):
    from oracle import su3 as o
    m = o.projectSU(_np(x))
    mt, vt = torch.from_numpy(m), torch.from_numpy(o.su3_to_vec(m))
    if want_matrix and want_vec:
        return mt, vt
    return mt if want_matrix else vt


def su3_project_vec(x, dtype=torch.float64):
    from oracle import su3 as o
    return torch.from_numpy(o.group_to_vec(_np(x))).to(dtype)


def su3_check(x):
    from oracle import su3 as o
    avg, mx = o.checkSU(_np(x))
    return torch.from_numpy(avg), torch.from_numpy(mx)


def su3_kinetic(p):
    from oracle import su3 as o
    return torch.from_numpy(o.kinetic_energy(_np(p)))


def su3_update_gauge(x, p, eps=1.0, mask=None, mask_complement=False, eps_mult=1.0):
    from oracle import su3 as o
    e = float(eps) * float(eps_mult)
    xn, pn = _np(x), _np(p).reshape(x.shape)
    if mask is None:
        return torch.from_numpy(o.update_gauge(xn, e * pn))
    m = _np(mask).astype(xn.real.dtype).reshape(1, *xn.shape[1:])
    if mask_complement:
        m = 1.0 - m
    return torch.from_numpy(m * xn + o.update_gauge((1.0 - m) * xn, e * pn))


def su3_vupdate(v, force, s, t, q, eps, sign):
    nb = v.shape[0]
    e = float(eps)
    vv, ff = v.reshape(nb, -1), force.reshape(nb, -1)
    z = torch.zeros(vv.shape, dtype=torch.float64)
    s, t, q = (z if a is None else a.to(torch.float64).reshape(nb, -1) for a in (s, t, q))
    kick = 0.5 * e * (ff * torch.exp(e * q) + t)
    if sign > 0:
        out, logdet = torch.exp(0.5 * e * s) * vv - kick, (0.5 * e * s).sum(1)
    else:
        out, logdet = torch.exp(-0.5 * e * s) * (vv + kick), -(0.5 * e * s).sum(1)
    return out.reshape(v.shape), logdet


def su3_hmc_trajectory(x, v, beta, eps, nlf):
    from oracle import su3 as o, dynamics as od
    xn, vn = _np(x), _np(v).reshape(x.shape)
    sp, _ = od.transition_kernel_hmc(od.SU3Ops, od.State(xn, vn, float(beta)), float(eps), int(nlf))
    en = torch.from_numpy(__import__('numpy').stack([o.kinetic_energy(vn), o.action(xn, float(beta)),
                                                     o.kinetic_energy(sp.v), o.action(sp.x, float(beta))], 1))
    return torch.from_numpy(sp.x), torch.from_numpy(sp.v), en


# ---- SU(3) adjoints: vjps of DIFFERENTIABLE torch restatements (independent of the kernels' closed forms) ----
def _t_project_su(x):
    """projectSU in torch, smooth at unitary input: the polar factor by Newton's iteration U <- (U + U^-+)/2
    (autograd through `inv` only; the closed-form eigen route is singular exactly where lattice links live), then
    the determinant phase exp(-i arg det / 3)"""
    u = x
    for _ in range(64):
        u = 0.5 * (u + torch.linalg.inv(u).mH)
    ph = torch.angle(torch.linalg.det(u)) / 3.0
    return u * torch.polar(torch.ones_like(ph), -ph)[..., None, None]


def _t_su3_to_vec(x):
    c = -2.0
    x00, x01, x02, x11, x12, x22 = x[..., 0, 0], x[..., 0, 1], x[..., 0, 2], x[..., 1, 1], x[..., 1, 2], x[..., 2, 2]
    return torch.stack([c * x01.imag, c * x01.real, x11.imag - x00.imag, c * x02.imag, c * x02.real, c * x12.imag,
                        c * x12.real, (2.0 * x22.imag - x11.imag - x00.imag) / 3.0 ** 0.5], -1)


def su3_project_bwd(x, gmat=None, gvec=None):
    with torch.enable_grad():
        lx = _leaf(x)
        y = _t_project_su(lx)
        outs, gs = [], []
        if gmat is not None:
            outs.append(y)
            gs.append(gmat)
        if gvec is not None:
            outs.append(_t_su3_to_vec(y))
            gs.append(gvec.to(torch.float64))
        return _vjp(outs, gs, [lx])[0]


def su3_vupdate_bwd(v, force, s, t, q, eps, sign, gv_out, glogdet):
    nb = v.shape[0]
    with torch.enable_grad():
        lv, lf = _leaf(v.reshape(nb, -1)), _leaf(force.reshape(nb, -1))
        ls, lt, lq = (None if a is None else _leaf(a.to(torch.float64).reshape(nb, -1)) for a in (s, t, q))
        e = torch.full((nb,), float(eps), dtype=torch.float64, requires_grad=True)
        z = torch.zeros(lv.shape, dtype=torch.float64)
        ss, tt, qq = (z if a is None else a for a in (ls, lt, lq))
        ee = e[:, None]
        kick = 0.5 * ee * (lf * torch.exp(ee * qq) + tt)
        if sign > 0:
            out, logdet = torch.exp(0.5 * ee * ss) * lv - kick, (0.5 * ee * ss).sum(1)
        else:
            out, logdet = torch.exp(-0.5 * ee * ss) * (lv + kick), -(0.5 * ee * ss).sum(1)
        gv, gf, gs, gt, gq, ge = _vjp([out, logdet], [gv_out.reshape(nb, -1), glogdet], [lv, lf, ls, lt, lq, e])
    return gv.reshape(v.shape), gf.reshape(v.shape), gs, gt, gq, ge


def su3_update_gauge_bwd(x, p, eps, mask, mask_complement, gx_out, eps_mult=1.0):
    nb = x.shape[0]
    with torch.enable_grad():
        lx, lp = _leaf(x), _leaf(p.reshape(x.shape))
        e = torch.full((nb,), float(eps) * float(eps_mult), dtype=torch.float64, requires_grad=True)
        ex = torch.linalg.matrix_exp(e.reshape(nb, *([1] * (x.dim() - 1))) * lp)
        if mask is None:
            out = ex @ lx
        else:
            m = mask.to(torch.float64).reshape(1, *x.shape[1:])
            if mask_complement:
                m = 1.0 - m
            out = m * lx + ex @ ((1.0 - m) * lx)
        gx, gp, ge = _vjp([out], [gx_out.reshape(x.shape)], [lx, lp, e])
    return gx, gp, ge, torch.zeros(1, dtype=torch.int32)


def su3_wilson_loops_bwd(x, gw):
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    lat = LatticeSU3(x.shape[0], list(x.shape[2:6]))
    with torch.enable_grad():
        lx = _leaf(x)
        w = torch.stack([lat._trace_plaquette(lx, u, v) for u in range(1, 4) for v in range(u)])
        return _vjp([w], [gw], [lx])[0]


def su3_rand_momentum(nb, dims, seed, offset, device, want_ke=False, offset_dev=None):
    """Gaussian traceless anti-Hermitian momenta from the oracle's generator, keyed like the kernel's Philox stream"""
    import numpy as np
    from oracle import su3 as o
    rng = np.random.default_rng([int(seed) & (2**63 - 1), int(offset)])
    p = torch.from_numpy(o.random_momentum(rng, (nb, 4, *[int(d) for d in dims], 3, 3)))
    return (p, su3_kinetic(p)) if want_ke else p


@contextlib.contextmanager
def su3_host_logic_on_cpu(monkeypatch):
    """as u1_host_logic_on_cpu, for the boundary-layout (unfused, no-grad) SU(3) path"""
    from l2hmc_b200 import ops
    from l2hmc_b200.dynamics.pytorch import dynamics as dmod
    from l2hmc_b200.network.pytorch import network as net
    from l2hmc_b200.group.su3.pytorch import group as g3
    for name in ('su3_plaq_sums', 'su3_wilson_loops', 'su3_force', 'su3_force_c1', 'su3_project', 'su3_project_vec',
                 'su3_kinetic', 'su3_check', 'su3_update_gauge', 'su3_vupdate', 'su3_hmc_trajectory', 'su3_rand_momentum', 'su3_action_grad_c1', 'su3_action_grad', 'su3_force_bwd', 'su3_project_bwd', 'su3_vupdate_bwd',
                 'su3_update_gauge_bwd', 'su3_wilson_loops_bwd', 'rowscale', 'accept_mix'):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, 'heads_supported', lambda hidden: False)         # tcgen05 heads: GPU tier only
    cpu = lambda: torch.device('cpu')  # noqa: E731
    monkeypatch.setattr(net, '_device', cpu)
    monkeypatch.setattr(g3, '_device', cpu)
    monkeypatch.setattr(dmod, 'torch', _TorchProxy())
    monkeypatch.setattr(torch.cuda, 'is_current_stream_capturing', lambda: False)
    yield


@contextlib.contextmanager
def u1_host_logic_on_cpu(monkeypatch):
    """patch l2hmc_b200 so that a U(1) `Dynamics` can be built and stepped on the CPU with the stand-ins above"""
    from l2hmc_b200 import ops
    from l2hmc_b200.dynamics.pytorch import dynamics as dmod
    from l2hmc_b200.network.pytorch import network as net
    from l2hmc_b200.group.u1.pytorch import group as gu1
    for name in ('u1_wilson_loops', 'u1_observables', 'u1_force', 'u1_kinetic', 'u1_compat_proj', 'u1_vupdate',
                 'u1_xupdate', 'u1_hmc_trajectory', 'accept_mix', 'u1_wilson_loops_bwd', 'u1_force_bwd', 'rowscale',
                 'u1_vupdate_bwd', 'u1_xupdate_bwd'):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, 'u1_heads_supported', lambda hidden: False)      # fused kernels: GPU tier only
    monkeypatch.setattr(ops, 'u1_input_supported', lambda units: False)
    cpu = lambda: torch.device('cpu')  # noqa: E731
    monkeypatch.setattr(net, '_device', cpu)
    monkeypatch.setattr(gu1, '_device', cpu)
    monkeypatch.setattr(dmod, 'torch', _TorchProxy())
    monkeypatch.setattr(torch.cuda, 'is_current_stream_capturing', lambda: False)   # Trainer.train_step asks
    yield


# ---------------------------------------------------------------------------
# stand-ins for the tensor-core GEMM and the conv-stack kernels (csrc/l2b_gemm.cu, l2b_conv.cu): plain CPU torch
# restatements of what include/l2b.h says they compute, so that the HOST wiring on top of them -- autograd.TCDense /
# ConvPeriodic / PoolAct, the deferred weight gradients, ConvStack's layer walk, ops.gemm_f32's segment lists --
# runs on the CPU tier.  bf16 rounding of operands and outputs is reproduced; accumulation is float32.
# ---------------------------------------------------------------------------
def _act_cpu(x, act):
    import torch.nn.functional as F
    return {None: lambda t: t, 'identity': lambda t: t, 'tanh': torch.tanh, 'relu': torch.relu, 'swish': F.silu,
            'leaky_relu': lambda t: F.leaky_relu(t, 0.01), 'elu': F.elu}[act](x)


def gemm_bf16(a, b, a_kmajor, b_kmajor, *, out=None, out_dtype=torch.bfloat16, bias=None, act=None, accumulate=False,
              splits=0, seg_inner=False):
    a_list = [a] if isinstance(a, torch.Tensor) else list(a)
    b_list = [b] if isinstance(b, torch.Tensor) else list(b)
    assert len(a_list) == len(b_list) and 1 <= len(a_list) <= 32
    acc = None
    for x, y in zip(a_list, b_list):
        assert x.dtype == torch.bfloat16 and y.dtype == torch.bfloat16
        A = x.float() if a_kmajor else x.float().t()
        B = y.float() if b_kmajor else y.float().t()
        K = min(A.shape[1], B.shape[1])
        assert (A.shape[1] + 7) // 8 == (B.shape[1] + 7) // 8, 'contraction lengths differ'
        assert float(A[:, K:].abs().sum()) == 0.0 and float(B[:, K:].abs().sum()) == 0.0   # only zero padding beyond K
        d = A[:, :K] @ B[:, :K].t()
        acc = d if acc is None else acc + d
    if bias is not None:
        acc = acc + bias.float().reshape(1, -1)[:, :acc.shape[1]]
    acc = _act_cpu(acc, act)
    if out is not None:
        out.copy_((out.float() + acc if accumulate else acc).to(out.dtype))
        return out
    return acc.to(out_dtype)


def split_bf16x3(x):
    assert x.dtype == torch.float32 and x.dim() == 2
    c8 = (x.shape[1] + 7) // 8 * 8
    xp = torch.nn.functional.pad(x, (0, c8 - x.shape[1]))
    a = xp.to(torch.bfloat16)
    r1 = xp - a.float()
    b = r1.to(torch.bfloat16)
    return torch.stack([a, b, (r1 - b.float()).to(torch.bfloat16)])


def _im2col_f32(x_nchw, n):
    size = n - 1
    xp = torch.cat([x_nchw[:, :, -size:, :], x_nchw, x_nchw[:, :, :size, :]], 2)
    xp = torch.cat([xp[:, :, :, -size:], xp, xp[:, :, :, :size]], 3)
    nb, C = x_nchw.shape[0], x_nchw.shape[1]
    return torch.nn.functional.unfold(xp, n).transpose(1, 2).reshape(-1, C * n * n)


def _tap_major(col, C, n):
    """columns (ci, kh, kw) -> (kh, kw, ci)"""
    return col.reshape(col.shape[0], C, n * n).transpose(1, 2).reshape(col.shape[0], C * n * n)


def conv_im2col(x, n, nchw, planes, tap_major=False):
    xf = (x if nchw else x.permute(0, 3, 1, 2)).float()
    col = _im2col_f32(xf, n)
    if tap_major:
        col = _tap_major(col, xf.shape[1], n)
    s3 = split_bf16x3(col.contiguous())
    return s3[:planes].contiguous()


def conv_col2im(dcol, like, n, nchw, tap_major=False):
    shape = like.shape if nchw else (like.shape[0], like.shape[3], like.shape[1], like.shape[2])
    K = shape[1] * n * n
    with torch.enable_grad():                  # called from inside a Function.backward (grad mode off)
        probe = torch.zeros(shape, dtype=torch.float32, requires_grad=True)
        col = _im2col_f32(probe, n)
        if tap_major:
            col = _tap_major(col, shape[1], n)
        (gx,) = torch.autograd.grad(col, probe, dcol.float()[:, :K])
    return gx if nchw else gx.permute(0, 2, 3, 1).contiguous()


def pool_act(x, pool, act):
    nb, H, W, C = x.shape
    xn = x.float().permute(0, 3, 1, 2)
    win = torch.nn.functional.unfold(xn, pool, stride=pool).reshape(nb, C, pool * pool, H // pool, W // pool)
    best, idx = win.max(2)
    y = _act_cpu(best, act).permute(0, 2, 3, 1).contiguous().to(x.dtype)
    return y, idx.permute(0, 2, 3, 1).contiguous().to(torch.uint8), (best.permute(0, 2, 3, 1).contiguous() if act == 'swish' else None)


def pool_act_bwd(gy, y, pre, idx, in_shape, pool, act):
    nb, H, W, C = in_shape
    yf = y.float()
    d = {None: lambda: torch.ones_like(yf), 'tanh': lambda: 1 - yf * yf, 'relu': lambda: (yf > 0).float(),
         'leaky_relu': lambda: torch.where(yf > 0, 1.0, 0.01), 'elu': lambda: torch.where(yf > 0, 1.0, yf + 1.0),
         'swish': lambda: torch.sigmoid(pre) * (1 + pre * (1 - torch.sigmoid(pre)))}[act]()
    g = gy.float() * d
    gx = torch.zeros(nb, H, W, C)
    PH, PW = H // pool, W // pool
    ii, jj = idx.long() // pool, idx.long() % pool
    b_, ph, pw, c = torch.meshgrid(torch.arange(nb), torch.arange(PH), torch.arange(PW), torch.arange(C), indexing='ij')
    gx[b_, ph * pool + ii, pw * pool + jj, c] = g
    return gx


@contextlib.contextmanager
def tensor_core_layers_on_cpu(monkeypatch):
    """the dense / conv layers' tensor-core path (TCDense, ConvPeriodic, PoolAct) on the CPU stand-ins above"""
    from l2hmc_b200 import ops
    from l2hmc_b200.network.pytorch import network as net
    for name in ('gemm_bf16', 'split_bf16x3', 'conv_im2col', 'conv_col2im', 'pool_act', 'pool_act_bwd'):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, '_need_cuda', lambda *ts: None)
    monkeypatch.setattr(net, '_tc_tensor_ok', lambda t: True)
    monkeypatch.setattr(net, '_device', lambda: torch.device('cpu'))
    monkeypatch.setattr(torch.cuda, 'is_current_stream_capturing', lambda: False)
    yield

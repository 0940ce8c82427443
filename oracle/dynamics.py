"""oracle/dynamics.py -- TEST INFRASTRUCTURE (the checker), never the product path.

Plain-numpy restatement of the reference's generalised leapfrog integrator
(`dynamics/pytorch/dynamics.py`): plain HMC (`leapfrog_hmc` :900-913,
`transition_kernel_hmc` :915-954) and the L2HMC forward/backward sweep
(`transition_kernel_fb` :956-1029, `_forward_lf/_backward_lf` :1187-1228,
`_update_v_{fwd,bwd}` :1266-1297, `_update_x_{fwd,bwd}` :1386-1477),
Metropolis-Hastings accept probability (:1065-1079) and the accept/reject mix
(:632-702).  Momenta, masks, accept uniforms and network weights are always
passed in explicitly: RNG streams cannot match across devices.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional, Sequence

import numpy as np

from . import su3 as _su3
from . import u1 as _u1
from . import network as _net


@dataclass
class State:
    x: np.ndarray
    v: np.ndarray
    beta: float


# ---------------------------------------------------------------------------
# group adaptors
# ---------------------------------------------------------------------------
class SU3Ops:
    name = 'SU3'
    action = staticmethod(_su3.action)
    grad_action = staticmethod(_su3.grad_action)
    kinetic_energy = staticmethod(_su3.kinetic_energy)
    update_gauge = staticmethod(_su3.update_gauge)
    compat_proj = staticmethod(_su3.projectSU)


class U1Ops:
    name = 'U1'
    action = staticmethod(_u1.action)
    grad_action = staticmethod(_u1.grad_action)
    kinetic_energy = staticmethod(_u1.kinetic_energy)
    update_gauge = staticmethod(_u1.update_gauge)
    compat_proj = staticmethod(_u1.compat_proj)


def hamiltonian(g, s: State):
    """dynamics.py:1479-1483"""
    return g.kinetic_energy(s.v) + g.action(s.x, s.beta)


def accept_prob(g, init: State, prop: State, sumlogdet):
    """exp(min(H0 - H1 + sumlogdet, 0))   (dynamics.py:1065-1079)"""
    dh = hamiltonian(g, init) - hamiltonian(g, prop) + sumlogdet
    return np.exp(np.minimum(dh, 0.0))


def accept_mix(acc, u, init: State, prop: State):
    """ma = (acc > u) as float32; out = ma*prop + mr*init  (dynamics.py:632-658,
    1081-1087).  Returns (x_out, v_out, ma), x/v flattened to [nb, -1]."""
    ma = (acc > u).astype(np.float32)
    mr = np.float32(1.0) - ma
    nb = init.x.shape[0]
    f = lambda a: a.reshape(nb, -1)  # noqa: E731
    xo = ma[:, None] * f(prop.x) + mr[:, None] * f(init.x)
    vo = ma[:, None] * f(prop.v) + mr[:, None] * f(init.v)
    return xo, vo, ma


# ---------------------------------------------------------------------------
# plain HMC
# ---------------------------------------------------------------------------
def leapfrog_hmc(g, s: State, eps: float, xshape=None) -> State:
    """dynamics.py:900-913: two un-merged half kicks around one drift.
    Like the reference, the returned x has v's shape (U(1): flat [nb, xdim])."""
    xshape = s.x.shape if xshape is None else xshape
    x_ = s.x.reshape(s.v.shape)
    eps = np.asarray(eps, dtype=s.v.real.dtype)
    half = np.asarray(0.5, dtype=s.v.real.dtype)
    f1 = g.grad_action(x_.reshape(xshape), s.beta).reshape(s.v.shape)
    v1 = s.v - half * eps * f1
    xp = g.update_gauge(x_, eps * v1)
    f2 = g.grad_action(xp.reshape(xshape), s.beta).reshape(s.v.shape)
    v2 = v1 - half * eps * f2
    return State(xp, v2, s.beta)


def transition_kernel_hmc(g, s: State, eps: float, nleapfrog: int):
    """dynamics.py:915-954 -> (proposed state, acc[nb])"""
    xshape = s.x.shape
    s_ = State(s.x, s.v, s.beta)
    for _ in range(nleapfrog):
        s_ = leapfrog_hmc(g, s_, eps, xshape)
    nb = s.x.shape[0]
    acc = accept_prob(g, s, State(s_.x.reshape(xshape), s_.v, s_.beta),
                      np.zeros(nb, dtype=s.v.real.dtype))
    return s_, acc


# ---------------------------------------------------------------------------
# L2HMC
# ---------------------------------------------------------------------------
@dataclass
class L2HMCSpec:
    """Everything `Dynamics` holds besides the state, made explicit."""
    group: str                                  # 'U1' | 'SU3'
    xshape: Sequence[int]                       # (nb, d, *lat[, 3, 3])
    nleapfrog: int
    xeps: Sequence[float]
    veps: Sequence[float]
    masks: Sequence[np.ndarray]                 # nlf x [1, xdim] float32
    state_dict: Optional[dict] = None           # reference Dynamics.state_dict() as numpy
    activation: str = 'tanh'
    use_batch_norm: bool = False
    use_ncp: bool = True
    use_split_xnets: bool = True
    use_separate_networks: bool = True
    nw_x: Sequence[float] = (1., 1., 1.)
    nw_v: Sequence[float] = (1., 1., 1.)
    conv: Optional[dict] = None
    nets: dict = field(default_factory=dict)

    @property
    def g(self):
        return SU3Ops if self.group == 'SU3' else U1Ops


def _sig_log(eps, dtype):
    """sigmoid(log(eps)) = eps / (1 + eps), computed the reference's way
    (dynamics.py:82-83,1270): 1 / (1 + exp(-log eps)) in the parameter's dtype
    (`torch.tensor(python_float)` -> torch's default dtype at construction)."""
    dt = np.dtype(dtype).type
    e = dt(eps)
    return dt(1.0) / (dt(1.0) + np.exp(-np.log(e)))


def _unflatten(spec: L2HMCSpec, a):
    return a.reshape(a.shape[0], *spec.xshape[1:])


def _net_prefix(spec: L2HMCSpec, kind: str, step: int, first: bool) -> str:
    """dynamics.py:1112-1136"""
    if kind == 'vnet':
        return f'vnet.{step}' if spec.use_separate_networks else 'vnet'
    if spec.use_separate_networks:
        if spec.use_split_xnets:
            return f'xnet.{step}.' + ('first' if first else 'second')
        return f'xnet.{step}'
    return 'xnet'


def call_vnet(spec: L2HMCSpec, step: int, x, force):
    """dynamics.py:1142-1159"""
    if spec.state_dict is None:           # dummy_network (network.py:69-77)
        z = np.zeros_like(x)
        return z, z, z
    if spec.group == 'SU3':
        x = _su3.group_to_vec(_unflatten(spec, x))
        force = _su3.group_to_vec(_unflatten(spec, force))
    sd = _net.sub_state_dict(spec.state_dict, _net_prefix(spec, 'vnet', step, False))
    conv_shape = None
    if spec.conv and spec.group == 'U1':
        conv_shape = (spec.xshape[1], *spec.xshape[2:4])
    return _net.leapfrog_layer(
        x, force, sd, activation=spec.activation, net_weight=spec.nw_v,
        use_batch_norm=spec.use_batch_norm,
        conv=spec.conv if spec.group == 'U1' else None, conv_in_shape=conv_shape)


def call_xnet(spec: L2HMCSpec, step: int, xm, v, first: bool):
    """dynamics.py:1161-1185 (U(1) only: the SU(3) x-update never calls xnet)"""
    if spec.state_dict is None:
        z = np.zeros_like(v.reshape(v.shape[0], -1))
        return z, z, z
    assert spec.group == 'U1'
    xin = _u1.group_to_vec(xm)            # [nb, 4, T, X]
    sd = _net.sub_state_dict(spec.state_dict, _net_prefix(spec, 'xnet', step, first))
    conv_shape = (spec.xshape[1] + 2, *spec.xshape[2:4]) if spec.conv else None
    return _net.leapfrog_layer(
        xin, v, sd, activation=spec.activation, net_weight=spec.nw_x,
        use_batch_norm=spec.use_batch_norm, conv=spec.conv, conv_in_shape=conv_shape)


def update_v(spec: L2HMCSpec, step: int, s: State, forward: bool):
    """dynamics.py:1266-1297"""
    g = spec.g
    force = g.grad_action(_unflatten(spec, s.x), s.beta)
    eps = _sig_log(spec.veps[step], s.v.real.dtype)
    sn, tn, qn = call_vnet(spec, step, s.x, force)
    vshape = s.v.shape
    force = force.reshape(vshape)
    if forward:
        logjac = eps * sn / 2.0
        logdet = logjac.reshape(logjac.shape[0], -1).sum(1)
        exp_s = np.exp(logjac).reshape(vshape)
        exp_q = np.exp(eps * qn).reshape(vshape)
        tn = tn.reshape(vshape)
        vf = exp_s * s.v - 0.5 * eps * (force * exp_q + tn)
    else:
        logjac = -eps * sn / 2.0
        logdet = logjac.reshape(logjac.shape[0], -1).sum(1)
        exp_s = np.exp(logjac).reshape(vshape)
        exp_q = np.exp(eps * qn).reshape(vshape)
        tn = tn.reshape(vshape)
        vf = exp_s * (s.v + 0.5 * eps * (force * exp_q + tn))
    return State(s.x, vf, s.beta), np.real(logdet)


def update_x(spec: L2HMCSpec, step: int, s: State, m, first: bool, forward: bool):
    """dynamics.py:1386-1477.  `m` is the [1, xdim] float32 mask."""
    eps = _sig_log(spec.xeps[step], s.v.real.dtype)
    m = _unflatten(spec, m)
    mb = np.ones_like(m) - m
    x = _unflatten(spec, s.x)
    xm_init = m * x
    nb = x.shape[0]
    if spec.group == 'U1':
        xf_ = x.reshape(nb, -1)
        v = s.v.reshape(nb, -1)
        sn, tn, qn = call_xnet(spec, step, xm_init, s.v, first)
        if forward:
            sn = eps * sn
            qn = eps * qn
            exp_s, exp_q = np.exp(sn), np.exp(qn)
            if spec.use_ncp:
                halfx = xf_ / 2.0
                _x = 2.0 * np.arctan(np.tan(halfx) * exp_s)
                xp = _unflatten(spec, _x + eps * (v * exp_q + tn))
                xn = xm_init + mb * xp
                cterm = np.cos(halfx) ** 2
                sterm = (exp_s * np.sin(halfx)) ** 2
                logdet_ = np.log(exp_s / (cterm + sterm))
                logdet = (mb.reshape(1, -1) * logdet_).sum(1)
            else:
                xp = xf_ * exp_s + eps * (v * exp_q + tn)
                xn = xm_init + mb * _unflatten(spec, xp)
                logdet = (mb.reshape(1, -1) * sn).sum(1)
        else:
            sn = (-eps) * sn
            qn = eps * qn
            exp_s, exp_q = np.exp(sn), np.exp(qn)
            if spec.use_ncp:
                halfx = xf_ / 2.0
                x1 = 2.0 * np.arctan(exp_s * np.tan(halfx))
                x2 = exp_s * eps * (v * exp_q + tn)
                xn = xm_init + mb * _unflatten(spec, x1 - x2)
                cterm = np.cos(halfx) ** 2
                sterm = (exp_s * np.sin(halfx)) ** 2
                logdet_ = np.log(exp_s / (cterm + sterm))
                logdet = (mb.reshape(1, -1) * logdet_).sum(1)
            else:
                xnew = exp_s * (xf_ - eps * (v * exp_q + tn))
                xn = xm_init + mb * _unflatten(spec, xnew)
                logdet = (mb.reshape(1, -1) * sn).sum(1)
        xn = _u1.compat_proj(xn)
    else:
        # SU(3): x' = m.x + exp(+-eps v) @ (mb.x); element-wise masks, logdet 0
        sign = 1.0 if forward else -1.0
        xn = xm_init + _su3.update_gauge(mb * x, (sign * eps) * s.v)
        logdet = np.zeros(nb)
    return State(xn, s.v, s.beta), np.real(logdet)


def forward_lf(spec: L2HMCSpec, step: int, s: State):
    """dynamics.py:1187-1207"""
    m = spec.masks[step]
    mb = np.ones_like(m) - m
    s, ld = update_v(spec, step, s, True)
    sld = ld
    s, ld = update_x(spec, step, s, m, True, True)
    sld = sld + ld
    s, ld = update_x(spec, step, s, mb, False, True)
    sld = sld + ld
    s, ld = update_v(spec, step, s, True)
    return s, sld + ld


def backward_lf(spec: L2HMCSpec, step: int, s: State):
    """dynamics.py:1209-1228"""
    step_r = spec.nleapfrog - step - 1
    m = spec.masks[step_r]
    mb = np.ones_like(m) - m
    s, ld = update_v(spec, step_r, s, False)
    sld = ld
    s, ld = update_x(spec, step_r, s, mb, False, False)
    sld = sld + ld
    s, ld = update_x(spec, step_r, s, m, True, False)
    sld = sld + ld
    s, ld = update_v(spec, step_r, s, False)
    return s, sld + ld


def transition_kernel_fb(spec: L2HMCSpec, s: State):
    """dynamics.py:956-1029 -> (proposed, acc, sumlogdet)"""
    nb = s.x.shape[0]
    sumlogdet = np.zeros(nb)
    s_ = State(s.x, s.v, s.beta)
    for step in range(spec.nleapfrog):
        s_, ld = forward_lf(spec, step, s_)
        sumlogdet = sumlogdet + ld
    s_ = State(s_.x, -s_.v, s_.beta)
    for step in range(spec.nleapfrog):
        s_, ld = backward_lf(spec, step, s_)
        sumlogdet = sumlogdet + ld
    acc = accept_prob(spec.g, State(_unflatten(spec, s.x), s.v, s.beta),
                      State(_unflatten(spec, s_.x), s_.v, s_.beta), sumlogdet)
    return s_, acc, sumlogdet


def transition_kernel(spec: L2HMCSpec, s: State, forward: bool):
    """dynamics.py:1031-1063, the un-merged kernel (merge_directions = False): nlf forward OR backward
    leapfrog layers.  NB the reference passes the states to `compute_accept_prob` SWAPPED
    (state_init = the final state, state_prop = the initial one, :1053-1057) -> (proposed, acc, sumlogdet)"""
    nb = s.x.shape[0]
    sumlogdet = np.zeros(nb)
    lf = forward_lf if forward else backward_lf
    s_ = State(s.x, s.v, s.beta)
    for step in range(spec.nleapfrog):
        s_, ld = lf(spec, step, s_)
        sumlogdet = sumlogdet + ld
    acc = accept_prob(spec.g, State(_unflatten(spec, s_.x), s_.v, s_.beta),
                      State(_unflatten(spec, s.x), s.v, s.beta), sumlogdet)
    return s_, acc, sumlogdet


def _metrics(spec: L2HMCSpec, s: State, logdet, extras=None, step=None):
    """get_metrics (dynamics.py:865-887)"""
    energy = hamiltonian(spec.g, State(_unflatten(spec, s.x), s.v, s.beta))
    m = {'energy': energy, 'logprob': energy - logdet, 'logdet': logdet}
    if extras:
        m.update(extras)
    if step is not None:
        m.update({'xeps': spec.xeps[step], 'veps': spec.veps[step]})
    return m


def _push(history: dict, metrics: dict):
    for k, v in metrics.items():
        history.setdefault(k, []).append(v)


def _stacked(history: dict) -> dict:
    return {k: (np.stack([np.asarray(e) for e in v]) if isinstance(v, list) else v) for k, v in history.items()}


def transition_kernel_fb_verbose(spec: L2HMCSpec, s: State):
    """transition_kernel_fb with config.verbose (dynamics.py:956-1029): the per-step metrics, recorded before the
    first layer and after each of the 2 nlf layers -> (proposed, history with [2 nlf + 1, nb] stacks)"""
    nb = s.x.shape[0]
    sumlogdet, sldf, sldb = np.zeros(nb), np.zeros(nb), np.zeros(nb)
    s_ = State(s.x, s.v, s.beta)
    h: dict = {}
    _push(h, _metrics(spec, s_, sumlogdet, {'sldf': sldf, 'sldb': sldb, 'sld': sumlogdet}, step=0))
    for step in range(spec.nleapfrog):
        s_, ld = forward_lf(spec, step, s_)
        sumlogdet = sumlogdet + ld
        sldf = sldf + ld
        _push(h, _metrics(spec, s_, sumlogdet, {'sldf': sldf, 'sldb': sldb, 'sld': sumlogdet}, step=step))
    s_ = State(s_.x, -s_.v, s_.beta)
    for step in range(spec.nleapfrog):
        s_, ld = backward_lf(spec, step, s_)
        sumlogdet = sumlogdet + ld
        sldb = sldb + ld
        _push(h, _metrics(spec, s_, sumlogdet, {'sldf': np.zeros(nb), 'sldb': sldb, 'sld': sumlogdet},
                          step=spec.nleapfrog - step - 1))
    acc = accept_prob(spec.g, State(_unflatten(spec, s.x), s.v, s.beta),
                      State(_unflatten(spec, s_.x), s_.v, s_.beta), sumlogdet)
    h.update({'acc': acc, 'sumlogdet': sumlogdet})
    return s_, _stacked(h)


def transition_kernel_verbose(spec: L2HMCSpec, s: State, forward: bool):
    """transition_kernel with config.verbose (dynamics.py:1031-1063)"""
    nb = s.x.shape[0]
    sumlogdet = np.zeros(nb)
    lf = forward_lf if forward else backward_lf
    s_ = State(s.x, s.v, s.beta)
    h: dict = {}
    _push(h, _metrics(spec, s_, sumlogdet))
    for step in range(spec.nleapfrog):
        s_, ld = lf(spec, step, s_)
        sumlogdet = sumlogdet + ld
        _push(h, _metrics(spec, s_, sumlogdet, step=step))
    acc = accept_prob(spec.g, State(_unflatten(spec, s_.x), s_.v, s_.beta),
                      State(_unflatten(spec, s.x), s.v, s.beta), sumlogdet)
    h.update({'acc': acc, 'sumlogdet': sumlogdet})
    return s_, _stacked(h)


def transition_kernel_hmc_verbose(spec: L2HMCSpec, s: State, eps: float, nleapfrog: int):
    """transition_kernel_hmc with config.verbose (dynamics.py:915-954): energies after every leapfrog step"""
    nb = s.x.shape[0]
    zeros = np.zeros(nb, dtype=s.v.real.dtype)
    xshape = s.x.shape
    s_ = State(s.x, s.v, s.beta)
    h: dict = {}
    _push(h, _metrics(spec, s_, zeros))
    for _ in range(nleapfrog):
        s_ = leapfrog_hmc(spec.g, s_, eps, xshape)
        _push(h, _metrics(spec, s_, zeros))
    acc = accept_prob(spec.g, s, State(s_.x.reshape(xshape), s_.v, s_.beta), zeros)
    h.update({'acc': acc, 'sumlogdet': zeros})
    return s_, _stacked(h)

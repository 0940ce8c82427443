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
    return eps.detach().to(like.dtype) if isinstance(eps, torch.Tensor) else torch.tensor(float(eps), dtype=like.dtype)


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


@contextlib.contextmanager
def u1_host_logic_on_cpu(monkeypatch):
    """patch l2hmc_b200 so that a U(1) `Dynamics` can be built and stepped on the CPU with the stand-ins above"""
    from l2hmc_b200 import ops
    from l2hmc_b200.dynamics.pytorch import dynamics as dmod
    from l2hmc_b200.network.pytorch import network as net
    from l2hmc_b200.group.u1.pytorch import group as gu1
    for name in ('u1_wilson_loops', 'u1_observables', 'u1_force', 'u1_kinetic', 'u1_compat_proj', 'u1_vupdate',
                 'u1_xupdate', 'u1_hmc_trajectory', 'accept_mix'):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(ops, 'u1_heads_supported', lambda hidden: False)      # fused kernels: GPU tier only
    monkeypatch.setattr(ops, 'u1_input_supported', lambda units: False)
    cpu = lambda: torch.device('cpu')  # noqa: E731
    monkeypatch.setattr(net, '_device', cpu)
    monkeypatch.setattr(gu1, '_device', cpu)
    monkeypatch.setattr(dmod, 'torch', _TorchProxy())
    yield

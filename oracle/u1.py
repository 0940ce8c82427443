"""oracle/u1.py -- TEST INFRASTRUCTURE (the checker), never the product path.

Plain-numpy CPU restatement of the reference's 2-D U(1) lattice + group
arithmetic; ``x[b, mu, t, x]`` real angles.  dtype follows the input (the
reference runs this group in float32 by default).  Paths are relative to
``/root/reference/src/l2hmc``.

The force is analytic instead of autograd (`lattice/u1/pytorch/lattice.py:102-117`);
both are checked against the reference in tests/ and frozen in tests/golden/.
"""
from __future__ import annotations

import numpy as np

PI = np.pi
TWO_PI = 2.0 * np.pi


def _c(x, val):
    """python scalar -> x's dtype (keeps float32 arithmetic float32, like torch)"""
    return np.asarray(val, dtype=x.dtype)


def compat_proj(x):
    """((x + pi) mod 2pi) - pi   (group/u1/pytorch/group.py:130-131)"""
    return np.mod(x + _c(x, PI), _c(x, TWO_PI)) - _c(x, PI)


def project_angle(x):
    """x - 2pi floor((x + pi) / 2pi)   (lattice/u1/pytorch/lattice.py:45-47)"""
    return x - _c(x, TWO_PI) * np.floor((x + _c(x, PI)) / _c(x, TWO_PI))


def wilson_loops(x):
    """xu + roll(xv,-1,T) - roll(xu,-1,X) - xv   (lattice.py:154-159)"""
    xu, xv = x[:, 0], x[:, 1]
    return xu + np.roll(xv, -1, axis=1) - np.roll(xu, -1, axis=2) - xv


def action(x, beta):
    """beta * sum(1 - cos P)   (lattice.py:80-86)"""
    w = wilson_loops(x)
    return _c(x, beta) * (_c(x, 1.0) - np.cos(w)).sum((1, 2), dtype=x.dtype)


def grad_action(x, beta):
    """dS/dx, analytic:  F0 = beta (sin P - roll(sin P,+1,X)),
                         F1 = beta (-sin P + roll(sin P,+1,T))
    (reference: autograd, lattice.py:102-117)"""
    s = np.sin(wilson_loops(x))
    b = _c(x, beta)
    f0 = b * (s - np.roll(s, 1, axis=2))
    f1 = b * (-s + np.roll(s, 1, axis=1))
    return np.stack([f0, f1], axis=1)


def plaqs(x):
    """mean cos P   (lattice.py:188-203)"""
    return np.cos(wilson_loops(x)).mean((1, 2), dtype=x.dtype)


def sin_charges(x):
    """sum sin P / 2pi   (lattice.py:221-224)"""
    return np.sin(wilson_loops(x)).sum((1, 2), dtype=x.dtype) / _c(x, TWO_PI)


def int_charges(x):
    """sum project_angle(P) / 2pi   (lattice.py:226-228)"""
    return project_angle(wilson_loops(x)).sum((1, 2), dtype=x.dtype) / _c(x, TWO_PI)


def kinetic_energy(v):
    """0.5 sum v^2   (group.py:164-165)"""
    nb = v.shape[0]
    return _c(v, 0.5) * (v.reshape(nb, -1) ** 2).sum(-1, dtype=v.dtype)


def update_gauge(x, p):
    """x + p   (group.py:99-100)"""
    return x + p


def group_to_vec(x):
    """cat(cos, sin) along dim 1   (group.py:86-88)"""
    return np.concatenate([np.cos(x), np.sin(x)], axis=1)


def wilson_loops4x4(x):
    """lattice.py:161-186 (note the trailing `.T` upstream: result is
    transposed to [X, T, nb])."""
    xu, xv = x[:, 0], x[:, 1]

    def r(a, st, sx):  # roll(shift_x over dims=2, shift_t over dims=1)
        return np.roll(np.roll(a, sx, axis=2), st, axis=1)
    w = (xu + r(xu, 0, -1) + r(xu, 0, -2) + r(xu, 0, -3) + r(xu, 0, -4)
         + r(xv, -1, -4) + r(xv, -2, -4) + r(xv, -3, -4)
         - r(xu, -4, -3) - r(xu, -4, -2) - r(xu, -4, -1)
         - r(xv, -4, 0) - r(xv, -3, 0) - r(xv, -2, 0) - r(xv, -1, 0) - xv)
    return w.T

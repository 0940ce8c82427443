"""oracle/su3.py -- TEST INFRASTRUCTURE (the checker), never the product path.

Plain-numpy CPU restatement of the reference's 4-D SU(3) lattice + group
arithmetic.  Layout is the reference's: ``x[b, mu, t, x, y, z, i, j]``
complex128 (`configs.py:501-507`).  All paths below are relative to
``/root/reference/src/l2hmc``.

Differences from the reference that are deliberate:
  * the force is analytic (staples) instead of autograd of the action
    (`lattice/su3/pytorch/lattice.py:299-308`); the two agree to ~2e-15
    (checked against the reference itself in tests/test_oracle_vs_reference.py
    and frozen in tests/golden/),
  * `expm` is an independent scaling-and-squaring Taylor series rather than
    ATen's `linalg_matrix_exp` (`group/su3/pytorch/group.py:45-50,88-90`).

PARITY PIN: the reference ships no tests or golden vectors for this path
(SURVEY.md section 4), so this oracle is pinned against outputs of the reference's
own PyTorch code run in this container: `oracle/make_golden.py` ->
`tests/golden/*.npz`, verified by `tests/test_oracle_golden.py`.
"""
from __future__ import annotations

import numpy as np

PI = np.pi
SQRT1by2 = np.sqrt(0.5)
SQRT1by3 = np.sqrt(1.0 / 3.0)
SQRT3 = np.sqrt(3.0)


# --------------------------------------------------------------------------
# elementary 3x3 helpers  (group/su3/pytorch/group.py:55-80)
# --------------------------------------------------------------------------
def adj(a: np.ndarray) -> np.ndarray:
    return np.conj(np.swapaxes(a, -1, -2))


def mul(a, b, adjoint_a: bool = False, adjoint_b: bool = False):
    """group.py:55-68"""
    if adjoint_a:
        a = adj(a)
    if adjoint_b:
        b = adj(b)
    return a @ b


def trace(a):
    """group.py:73-75"""
    return np.trace(a, axis1=-2, axis2=-1)


def det3(a):
    """closed-form 3x3 determinant (reference: torch `.det()`, LU based)."""
    return (a[..., 0, 0] * (a[..., 1, 1] * a[..., 2, 2] - a[..., 1, 2] * a[..., 2, 1])
            - a[..., 0, 1] * (a[..., 1, 0] * a[..., 2, 2] - a[..., 1, 2] * a[..., 2, 0])
            + a[..., 0, 2] * (a[..., 1, 0] * a[..., 2, 1] - a[..., 1, 1] * a[..., 2, 0]))


def eye_like(a):
    return np.broadcast_to(np.eye(3, dtype=a.dtype), a.shape)


def norm2(a):
    """utils.py:157-168 with default axes: sum_ij |a_ij|^2"""
    return (np.abs(a) ** 2).sum((-2, -1))


# --------------------------------------------------------------------------
# algebra  (group.py:92-103,125-126; utils.py:171-195,394-445)
# --------------------------------------------------------------------------
def projectTAH(x):
    """R = (X - X^+)/2 - tr(.)/3   (group.py:92-103)"""
    r = 0.5 * (x - adj(x))
    d = trace(r) / 3.0
    return r - d[..., None, None] * np.eye(3)


def kinetic_energy(p):
    """0.5 * sum_links (|P|_F^2 - 8)   (group.py:125-126)"""
    nb = p.shape[0]
    return 0.5 * (norm2(p) - 8.0).reshape(nb, -1).sum(1)


def tah_from_normals(n8: np.ndarray) -> np.ndarray:
    """`randTAH3` with the eight N(0,1) draws given explicitly.

    n8[..., k], k = (r3, r8, r01, r02, r12, i01, i02, i12) in the ORDER the
    reference draws them (utils.py:171-195).  Returns [..., 3, 3]."""
    r3 = SQRT1by2 * n8[..., 0]
    r8 = SQRT1by2 * SQRT1by3 * n8[..., 1]
    r01, r02, r12 = (SQRT1by2 * n8[..., k] for k in (2, 3, 4))
    i01, i02, i12 = (SQRT1by2 * n8[..., k] for k in (5, 6, 7))
    m = np.zeros(n8.shape[:-1] + (3, 3), dtype=np.complex128)
    m[..., 0, 0] = 1j * (r8 + r3)
    m[..., 1, 1] = 1j * (r8 - r3)
    m[..., 2, 2] = 1j * (-2.0 * r8)
    m[..., 0, 1] = r01 + 1j * i01
    m[..., 1, 0] = -r01 + 1j * i01
    m[..., 0, 2] = r02 + 1j * i02
    m[..., 2, 0] = -r02 + 1j * i02
    m[..., 1, 2] = r12 + 1j * i12
    m[..., 2, 1] = -r12 + 1j * i12
    return m


def su3_to_vec(x):
    """utils.py:394-420"""
    c = -2.0
    x00, x01, x02 = x[..., 0, 0], x[..., 0, 1], x[..., 0, 2]
    x11, x12, x22 = x[..., 1, 1], x[..., 1, 2], x[..., 2, 2]
    return np.stack([
        c * x01.imag, c * x01.real, x11.imag - x00.imag,
        c * x02.imag, c * x02.real, c * x12.imag, c * x12.real,
        SQRT1by3 * (2.0 * x22.imag - x11.imag - x00.imag),
    ], axis=-1)


def vec_to_su3(v):
    """utils.py:423-445"""
    c = -0.5
    x01 = c * (v[..., 1] + 1j * v[..., 0])
    x02 = c * (v[..., 4] + 1j * v[..., 3])
    x12 = c * (v[..., 6] + 1j * v[..., 5])
    x2i = SQRT1by3 * v[..., 7]
    x0i = c * (x2i + v[..., 2])
    x1i = c * (x2i - v[..., 2])
    m = np.zeros(v.shape[:-1] + (3, 3), dtype=np.complex128)
    # note the reference stacks COLUMNS (stack(..., -1) of stack(..., -1))
    m[..., 0, 0] = 1j * x0i
    m[..., 1, 1] = 1j * x1i
    m[..., 2, 2] = 1j * x2i
    m[..., 0, 1] = x01
    m[..., 0, 2] = x02
    m[..., 1, 2] = x12
    m[..., 1, 0] = -np.conj(x01)
    m[..., 2, 0] = -np.conj(x02)
    m[..., 2, 1] = -np.conj(x12)
    return m


# --------------------------------------------------------------------------
# projectSU  (utils.py:227-346)
# --------------------------------------------------------------------------
def eigs3x3(tr, p2, det):
    """utils.py:227-283, clamps included."""
    tr3 = tr / 3.0
    p23 = p2 / 3.0
    tr32 = tr3 * tr3
    q = np.abs(0.5 * (p23 - tr32))
    r = 0.25 * tr3 * (5.0 * tr32 - p2) - 0.5 * det
    sq = np.sqrt(q)
    sq3 = q * sq
    with np.errstate(divide='ignore', invalid='ignore'):
        isq3 = 1.0 / sq3
    isq3c = np.minimum(3e38, np.maximum(-3e38, isq3))
    rsq3c = r * isq3c
    rsq3 = np.minimum(1.0, np.maximum(-1.0, rsq3c))
    rsq3 = np.clip(rsq3, -1.0 + 1e-12, 1.0 - 1e-12)
    t = np.arccos(rsq3) / 3.0
    st, ct = np.sin(t), np.cos(t)
    sqc = sq * ct
    sqs = SQRT3 * sq * st
    ll = tr3 + sqc
    return tr3 - 2.0 * sqc, ll + sqs, ll - sqs


def rsqrtPHM3f(tr, p2, det):
    """utils.py:286-317"""
    e0, e1, e2 = eigs3x3(tr, p2, det)
    se0, se1, se2 = (np.sqrt(np.abs(e)) for e in (e0, e1, e2))
    u = se0 + se1 + se2
    w = se0 * se1 * se2
    d = w * (se0 + se1) * (se0 + se2) * (se1 + se2)
    di = 1.0 / d
    c0 = di * (w * u * u + e0 * se0 * (e1 + e2) + e1 * se1 * (e0 + e2)
               + e2 * se2 * (e0 + e1))
    c1 = -(tr * u + w) * di
    c2 = u * di
    return c0, c1, c2


def rsqrtPHM3(x):
    """utils.py:320-329"""
    tr = trace(x).real
    x2 = x @ x
    p2 = trace(x2).real
    det = det3(x).real
    c0, c1, c2 = rsqrtPHM3f(tr, p2, det)
    return (c0[..., None, None] * np.eye(3) + c1[..., None, None] * x
            + c2[..., None, None] * x2)


def projectU(x):
    """x (x^+ x)^{-1/2}   (utils.py:332-338)"""
    return x @ rsqrtPHM3(adj(x) @ x)


def projectSU(x):
    """utils.py:341-346"""
    m = projectU(x)
    d = det3(m)
    p = -np.arctan2(d.imag, d.real) / 3.0
    return m * (np.cos(p) + 1j * np.sin(p))[..., None, None]


def checkSU(x):
    """utils.py:376-391 -> (avg, max) per chain"""
    nb = x.shape[0]
    d = norm2(adj(x) @ x - np.eye(3))
    d = d + np.abs(-1.0 + det3(x)) ** 2
    d = d.reshape(nb, -1)
    c = 2.0 * (3 * 3 + 1)
    return np.sqrt(d.mean(1) / c), np.sqrt(d.max(1) / c)


def group_to_vec(x):
    """group.py:138-147"""
    return su3_to_vec(projectSU(x))


def random_su3(rng: np.random.Generator, shape):
    """group.py:113-119 with numpy RNG (stream parity is impossible)."""
    r = rng.standard_normal(tuple(shape))
    i = rng.standard_normal(tuple(shape))
    return projectSU(r + 1j * i)


def random_momentum(rng: np.random.Generator, shape):
    """group.py:121-123 / utils.py:171-195; `shape` is the full link shape."""
    n8 = rng.standard_normal(tuple(shape[:-2]) + (8,))
    return tah_from_normals(n8)


# --------------------------------------------------------------------------
# matrix exponential and link update  (group.py:45-50,88-90)
# --------------------------------------------------------------------------
def expm(a: np.ndarray, order: int = 18) -> np.ndarray:
    """Batched 3x3 matrix exponential: scaling-and-squaring + Horner Taylor.

    Scale so that ||A/2^s||_F <= 0.5; order-18 Taylor then has a truncation
    error < 0.5^19/19! ~ 1.6e-23."""
    a = np.asarray(a, dtype=np.complex128)
    nrm = np.sqrt(norm2(a))
    with np.errstate(divide='ignore'):
        s = np.where(nrm > 0.5, np.ceil(np.log2(np.maximum(nrm, 1e-300) / 0.5)), 0.0)
    s = s.astype(np.int64)
    smax = int(s.max()) if s.size else 0
    a = a * (0.5 ** s)[..., None, None]
    eye = np.eye(3, dtype=np.complex128)
    x = eye + a / order
    for i in range(order - 1, 0, -1):
        x = eye + (a @ x) / i
    for k in range(smax):
        x = np.where((s > k)[..., None, None], x @ x, x)
    return x


def update_gauge(x, p):
    """exp(P) U   (group.py:45-50)"""
    return expm(p) @ x


# --------------------------------------------------------------------------
# lattice  (lattice/su3/pytorch/lattice.py)
# --------------------------------------------------------------------------
def _shift(a, mu, sign=+1):
    """a(n + sign*mu_hat) on an array whose site axes are 1..4 (chain axis 0)."""
    return np.roll(a, -sign, axis=mu + 1)


def wilson_loops(x):
    """`_wilson_loops` (lattice.py:157-199), c1 == 0: ps[6, nb, T, X, Y, Z]."""
    ps = []
    for u in range(1, 4):
        for v in range(0, u):
            xu, xv = x[:, u], x[:, v]
            yuv = xu @ _shift(xv, u)
            yvu = xv @ _shift(xu, v)
            ps.append(trace(yuv @ adj(yvu)))
    return np.stack(ps)


def plaq_sums(x):
    """(sum Re tr P, sum Im tr P) per chain"""
    ps = wilson_loops(x)
    nb = x.shape[0]
    psr = ps.real.reshape(6, nb, -1).sum(2).sum(0)
    psi = ps.imag.reshape(6, nb, -1).sum(2).sum(0)
    return psr, psi


def rect_traces(x):
    """traces of the 2x1 / 1x2 rectangles, rs[12, nb, T, X, Y, Z] (`_wilson_loops` with
    needs_rect, lattice.py:180-196)"""
    rs = []
    for u in range(1, 4):
        for v in range(0, u):
            xu, xv = x[:, u], x[:, v]
            yuv = xu @ _shift(xv, u)
            yvu = xv @ _shift(xu, v)
            yu, yv = _shift(xu, v), _shift(xv, u)
            uu, ur = adj(xv) @ yuv, adj(xu) @ yvu
            ul, ud = yuv @ adj(yu), yvu @ adj(yv)
            rs.append(trace(ur @ adj(_shift(ul, u))))
            rs.append(trace(uu @ adj(_shift(ud, v))))
    return np.stack(rs)


def action(x, beta, c1: float = 0.0):
    """lattice.py:252-269: -(beta (1 - 8 c1) / 3) sum Re tr P - (beta c1 / 3) sum Re tr R"""
    psr, _ = plaq_sums(x)
    s = beta * (1.0 - 8.0 * c1) * psr
    if c1 != 0.0:
        rs = rect_traces(x)
        s = s + beta * c1 * rs.real.reshape(12, x.shape[0], -1).sum(2).sum(0)
    return s * (-1.0 / 3.0)


def volume_of(x):
    return int(np.prod(x.shape[2:6]))


def plaqs(x):
    """lattice.py:201-211"""
    return plaq_sums(x)[0] / (6 * 3 * volume_of(x))


def int_charges(x):
    """lattice.py:232-235 (marked TODO upstream; not an integer)"""
    return plaq_sums(x)[1] / (32 * PI ** 2)


def sin_charges(x):
    """lattice.py:237-240"""
    return plaq_sums(x)[1] / (6 * 3 * volume_of(x))


def staples(x):
    """A_mu(n) = sum_{nu != mu} [ U_nu(n+mu) U_mu(n+nu)^+ U_nu(n)^+
                                 + U_nu(n+mu-nu)^+ U_mu(n-nu)^+ U_nu(n-nu) ]"""
    out = np.zeros_like(x)
    for mu in range(4):
        a = np.zeros_like(x[:, mu])
        for nu in range(4):
            if nu == mu:
                continue
            umu, unu = x[:, mu], x[:, nu]
            fwd = _shift(unu, mu) @ adj(_shift(umu, nu)) @ adj(unu)
            unu_mn = _shift(unu, nu, -1)                # U_nu(n-nu)
            umu_mn = _shift(umu, nu, -1)                # U_mu(n-nu)
            unu_pm = _shift(unu_mn, mu)                 # U_nu(n+mu-nu)
            bwd = adj(unu_pm) @ adj(umu_mn) @ unu_mn
            a = a + fwd + bwd
        out[:, mu] = a
    return out


def grad_action(x, beta):
    """F = projectTAH(dS/dU U^+) == (beta/3) TAH(U A)   (lattice.py:299-308)"""
    return (beta / 3.0) * projectTAH(x @ staples(x))

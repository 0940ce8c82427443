"""CPU tier: the __host__ __device__ bodies of the CUDA kernels
(l2hmc_b200/csrc/l2b_su3_math.cuh, l2b_su3_site.cuh), compiled for the host by
tests/hostemu, agree with the golden vectors of the reference.  This checks the
kernels' arithmetic and lattice index math without a GPU; launch geometry,
shared-memory staging and reductions are covered by the `-m gpu` tier."""
import ctypes
import subprocess
from pathlib import Path

import numpy as np
import pytest

from oracle import su3 as osu3

HERE = Path(__file__).resolve().parent / 'hostemu'
P = ctypes.c_void_p


@pytest.fixture(scope='module')
def emu():
    so = HERE / 'libhostemu.so'
    src = HERE / 'hostemu.cpp'
    deps = [src] + list((HERE.parents[1] / 'l2hmc_b200' / 'csrc').glob('*.cuh'))
    if not so.exists() or any(d.stat().st_mtime > so.stat().st_mtime for d in deps):
        subprocess.check_call(['g++', '-O2', '-std=c++17', '-shared', '-fPIC', '-o', str(so), str(src)])
    return ctypes.CDLL(str(so))


@pytest.fixture(scope='module')
def g(golden_dir):
    return np.load(golden_dir / 'su3_f64.npz')


def ptr(a):
    return None if a is None else a.ctypes.data_as(P)


def maxdiff(a, b):
    return float(np.max(np.abs(a - b)))


def unary(emu, mode, a, scale=1.0, vec=False):
    a = np.ascontiguousarray(a)
    out = np.empty_like(a)
    v8 = np.empty(a.shape[:-2] + (8,)) if vec else None
    emu.emu_unary(ctypes.c_int(mode), ptr(a), ctypes.c_double(scale), ptr(out), ptr(v8), ctypes.c_size_t(a.size // 9))
    return (out, v8) if vec else out


def test_link_local_math(emu, g):
    assert maxdiff(unary(emu, 0, g['v'], 0.25), g['expv']) < 1e-14
    assert maxdiff(unary(emu, 0, g['y'], 1.0), g['expy']) < 1e-11 * np.abs(g['expy']).max()
    assert maxdiff(unary(emu, 1, g['y']), g['tah_y']) < 1e-15
    assert maxdiff(unary(emu, 2, g['y']), g['projsu_y']) < 1e-10
    _, vx = unary(emu, 2, g['x'], vec=True)
    assert maxdiff(vx, g['vec_x']) < 1e-13
    m = np.empty_like(g['x'])
    emu.emu_from_vec(ptr(np.ascontiguousarray(g['vec_x'])), ptr(m), ctypes.c_size_t(m.size // 9))
    assert maxdiff(m, g['vec2su3']) < 1e-15
    rng = np.random.default_rng(0)
    n8 = rng.standard_normal((64, 8))
    t = np.empty((64, 3, 3), dtype=np.complex128)
    emu.emu_tah_from_normals(ptr(n8), ptr(t), ctypes.c_size_t(64))
    assert maxdiff(t, osu3.tah_from_normals(n8)) == 0.0
    y = np.ascontiguousarray(g['y'])
    d = np.empty(y.shape[:-2])
    emu.emu_check(ptr(y), ptr(d), ctypes.c_size_t(y.size // 9))
    nb = y.shape[0]
    assert np.allclose(np.sqrt(d.reshape(nb, -1).mean(1) / 20), g['checksu_avg'], rtol=1e-13)


def test_exp_large_norm_uses_scaling(emu):
    rng = np.random.default_rng(3)
    a = (rng.standard_normal((32, 3, 3)) + 1j * rng.standard_normal((32, 3, 3))) * 6.0
    ref = osu3.expm(a)
    out = unary(emu, 0, a)
    assert np.max(np.abs(out - ref) / np.abs(ref).max()) < 1e-11


def test_exp_specialised_to_the_algebra(emu, g):
    """mat_exp_alg (the HMC drift's exponential): su(3) arguments take the specialised Cayley-Hamilton
    series (real c, imaginary d, Hermitian A^2, 1/n! table), everything else falls back to mat_exp"""
    rng = np.random.default_rng(11)
    p = osu3.random_momentum(rng, (4096, 3, 3))              # ||P||_F^2 ~ 8
    for scale in (1e-3, 0.02, 0.05, 0.1, 0.17, 0.3, 0.35):   # leapfrog arguments eps * P, all Taylor-order buckets
        a = scale * p
        ref = osu3.expm(a)
        got = unary(emu, 4, a)
        assert maxdiff(got, ref) < 1e-15, scale
        assert maxdiff(got, unary(emu, 0, a)) < 1e-15        # vs the general series: rounding only
        # unitary to rounding
        assert maxdiff(got @ got.conj().swapaxes(-1, -2), np.broadcast_to(np.eye(3), got.shape)) < 2e-15
    assert maxdiff(unary(emu, 4, g['v'], 0.25), g['expv']) < 1e-14
    # not in the algebra (generic complex, Hermitian part, trace) or too large: same bits as mat_exp
    gen = (rng.standard_normal((64, 3, 3)) + 1j * rng.standard_normal((64, 3, 3))) * 0.2
    assert np.array_equal(unary(emu, 4, gen), unary(emu, 0, gen))
    tr = 0.1 * p[:64] + 1e-9j * np.eye(3)
    assert np.array_equal(unary(emu, 4, tr), unary(emu, 0, tr))
    big = 2.0 * p[:64]
    assert np.array_equal(unary(emu, 4, big), unary(emu, 0, big))


def test_stencil_bodies(emu, g):
    x = np.ascontiguousarray(g['x'])
    nb, beta = x.shape[0], float(g['beta'])
    dims = (ctypes.c_int * 4)(*[int(s) for s in g['shape']])
    f = np.empty_like(x)
    retr = np.empty(nb)
    emu.emu_force(ptr(x), ctypes.c_double(beta), ptr(f), ptr(retr), ctypes.c_int(nb), dims)
    assert maxdiff(f, g['force']) < 1e-13
    assert np.allclose(-beta / 3 * retr / 4, g['action'], rtol=1e-12, atol=1e-12)
    wl = np.empty((6, nb) + x.shape[2:6], dtype=np.complex128)
    emu.emu_wloops(ptr(x), ptr(wl), ctypes.c_int(nb), dims)
    assert maxdiff(wl, g['wloops']) < 1e-13
    # the hooked variant the default kick kernels use: bit-identical, hook runs exactly once per link
    for hook_at in (0, 2, 3, 12, 13):      # 12 / 13: the row-streaming low-register form, hook at 2 / 3
        fh = np.empty_like(x)
        hooks = np.zeros(nb, dtype=np.int64)
        emu.emu_force_hook(ptr(x), ctypes.c_double(beta), ptr(fh), ptr(hooks), ctypes.c_int(hook_at), ctypes.c_int(nb),
                           dims)
        assert np.array_equal(fh, f) and np.all(hooks == x[0].size // 9)


@pytest.mark.parametrize('key,xk,vk', [('hmc1', 'x', 'v'), ('hmc4', 'x', 'v'), ('hmcw', 'xw', 'vw')])
def test_trajectory_kernel_sequence(emu, g, key, xk, vk):
    """merged kicks == the reference's two half kicks, to rounding"""
    x, v = np.ascontiguousarray(g[xk]), np.ascontiguousarray(g[vk])
    nb, beta = x.shape[0], float(g['beta'])
    nlf = int(g['hmcw_nlf']) if key == 'hmcw' else int(key[3:])
    dims = (ctypes.c_int * 4)(*[int(s) for s in g['shape']])
    xo, vo, en = np.empty_like(x), np.empty_like(x), np.empty((nb, 4))
    emu.emu_hmc(ptr(x), ptr(v), ctypes.c_double(beta), ctypes.c_double(float(g[f'{key}_eps'])), ctypes.c_int(nlf),
                ptr(xo), ptr(vo), ptr(en), ctypes.c_int(nb), dims)
    assert maxdiff(xo, g[f'{key}_x']) < 1e-13
    assert maxdiff(vo, g[f'{key}_v']) < 1e-13
    h0, h1 = en[:, 0] + en[:, 1], en[:, 2] + en[:, 3]
    assert np.allclose(h0, g[f'{key}_h0'], rtol=1e-12, atol=1e-11)
    assert np.allclose(h1, g[f'{key}_h1'], rtol=1e-12, atol=1e-11)
    acc = np.exp(np.minimum(h0 - h1, 0.0))
    assert maxdiff(acc, g[f'{key}_acc']) < 1e-11


def test_project_su_adjoint_matches_reference_autograd(emu, golden_dir):
    """closed-form adjoint of projectSU / group_to_vec (polar factor + Sylvester solve,
    l2b_su3_math.cuh project_su_adjoint) == the reference's autograd VJPs"""
    ga = np.load(golden_dir / 'su3_adjoint_f64.npz')
    x = np.ascontiguousarray(ga['x'])
    n = x.shape[0]
    for gmat, gvec, want in ((ga['gmat'], None, ga['gx_mat']), (None, ga['gvec'], ga['gx_vec'])):
        gx = np.empty_like(x)
        emu.emu_project_bwd(ptr(x), ptr(None if gmat is None else np.ascontiguousarray(gmat)),
                            ptr(None if gvec is None else np.ascontiguousarray(gvec)), ptr(gx), ctypes.c_size_t(n))
        # the map is ill conditioned for the anti-Hermitian block (|g| up to 1e3): relative per link
        scale = np.maximum(1.0, np.abs(want).reshape(n, -1).max(1))
        err = np.abs(gx - want).reshape(n, -1).max(1) / scale
        assert err.max() < 1e-8, err.max()
        assert np.median(err) < 1e-12


def test_wilson_loops_adjoint_matches_reference_autograd(emu, golden_dir):
    """weighted-staple adjoint of the per-site Wilson loops (wloops_adjoint_link) == the
    reference's autograd through lattice.py:157-199"""
    ga = np.load(golden_dir / 'su3_adjoint_f64.npz')
    x, gw, want = (np.ascontiguousarray(ga[k]) for k in ('wl_x', 'wl_gw', 'wl_gx'))
    dims = (ctypes.c_int * 4)(*[int(v) for v in ga['wl_shape']])
    gx = np.empty_like(x)
    emu.emu_wloops_bwd(ptr(x), ptr(gw), ptr(gx), ctypes.c_int(x.shape[0]), dims)
    assert maxdiff(gx, want) < 1e-13 * max(1.0, np.abs(want).max())


def test_exp_adjoint_matches_torch_autograd_at_all_norms(emu):
    """mat_exp_adjoint (Taylor adjoint on Cayley-Hamilton coefficients, term count chosen from the
    norm) against torch's autograd of matrix_exp, at the thresholds of the term-count table"""
    import torch
    rng = np.random.default_rng(3)
    for rho in (0.03, 0.0999, 0.1001, 0.2999, 0.3001, 0.4999, 0.5001, 0.999, 1.001, 1.999, 2.001, 2.95):
        a = rng.standard_normal((40, 3, 3)) + 1j * rng.standard_normal((40, 3, 3))
        a = np.ascontiguousarray(a / np.linalg.norm(a, axis=(1, 2), keepdims=True) * rho)
        g = np.ascontiguousarray(rng.standard_normal((40, 3, 3)) + 1j * rng.standard_normal((40, 3, 3)))
        ga = np.empty_like(a)
        emu.emu_exp_adjoint(ptr(a), ptr(g), ptr(ga), ctypes.c_size_t(40))
        at = torch.from_numpy(a).requires_grad_(True)
        want, = torch.autograd.grad(torch.linalg.matrix_exp(at), at, grad_outputs=torch.from_numpy(g))
        assert maxdiff(ga, want.numpy()) < 2e-14 * max(1.0, np.abs(want.numpy()).max()), rho


def test_rectangle_force_body_matches_reference_c1(emu, golden_dir):
    """rect_staples / link_times_improved_staples (the c1 != 0 force kernel's body) against the
    reference's own c1 = -0.331 run (autograd force, lattice.py:96-112,252-269,299-308) on a
    2 x 4 x 3 x 2 lattice: extents 2 and 3 exercise the wrap of the two-site hops"""
    g = np.load(golden_dir / 'su3_c1_f64.npz')
    x = np.ascontiguousarray(g['x'])
    nb, beta, c1 = x.shape[0], float(g['beta']), float(g['c1'])
    dims = (ctypes.c_int * 4)(*[int(s) for s in g['shape']])
    f = np.empty_like(x)
    sums = np.empty((nb, 2))
    emu.emu_force_c1(ptr(x), ctypes.c_double(beta), ctypes.c_double(c1), ptr(f), ptr(sums), ctypes.c_int(nb), dims)
    assert maxdiff(f, g['force']) < 1e-12 * max(1.0, np.abs(g['force']).max())
    rs = g['rects'].real.reshape(12, nb, -1).sum(2).sum(0)
    assert np.allclose(sums[:, 1], rs, rtol=1e-12, atol=1e-11)
    assert np.allclose(sums[:, 0], osu3.plaq_sums(x)[0], rtol=1e-12, atol=1e-11)
    s = -(beta / 3.0) * ((1 - 8 * c1) * sums[:, 0] + c1 * sums[:, 1])
    assert np.allclose(s, g['action'], rtol=1e-12, atol=1e-11)
    # c1 = 0 reduces to the plaquette force
    f0 = np.empty_like(x)
    emu.emu_force_c1(ptr(x), ctypes.c_double(beta), ctypes.c_double(0.0), ptr(f0), ptr(sums), ctypes.c_int(nb), dims)
    fp = np.empty_like(x)
    emu.emu_force(ptr(x), ctypes.c_double(beta), ptr(fp), None, ctypes.c_int(nb), dims)
    assert maxdiff(f0, fp) < 1e-14


def _gauge_transform(x, gt):
    """U_mu(n) -> g(n) U_mu(n) g(n + mu)^+"""
    out = np.empty_like(x)
    for mu in range(4):
        out[:, mu] = gt @ x[:, mu] @ osu3.adj(np.roll(gt, -1, axis=1 + mu))
    return out


@pytest.mark.parametrize('c1', [0.0, -0.331])
def test_gauge_covariance_of_the_force_bodies(emu, c1):
    """domain property, independent of any golden: under a local gauge transformation the action is invariant
    and the force rotates, F_mu(n) -> g(n) F_mu(n) g(n)^+; holds for the plaquette and the rectangle staples"""
    rng = np.random.default_rng(17)
    shape, nb, beta = (4, 3, 2, 5), 2, 5.7
    x = np.ascontiguousarray(osu3.random_su3(rng, (nb, 4, *shape, 3, 3)))
    gt = osu3.random_su3(rng, (nb, *shape, 3, 3))
    xg = np.ascontiguousarray(_gauge_transform(x, gt))
    dims = (ctypes.c_int * 4)(*shape)
    f, fg = np.empty_like(x), np.empty_like(x)
    s, sg = np.empty((nb, 2)), np.empty((nb, 2))
    for xx, ff, ss in ((x, f, s), (xg, fg, sg)):
        emu.emu_force_c1(ptr(xx), ctypes.c_double(beta), ctypes.c_double(c1), ptr(ff), ptr(ss), ctypes.c_int(nb), dims)
    assert np.allclose(s, sg, rtol=1e-12, atol=1e-10)
    want = np.empty_like(f)
    for mu in range(4):
        want[:, mu] = gt @ f[:, mu] @ osu3.adj(gt)
    assert maxdiff(fg, want) < 1e-12 * max(1.0, np.abs(f).max())
    # the force is traceless anti-Hermitian
    assert maxdiff(f, -osu3.adj(f)) < 1e-13 and np.abs(osu3.trace(f)).max() < 1e-13


def test_trajectory_body_is_reversible_and_stays_in_the_group(emu):
    """leapfrog is time-reversible: integrating forward, flipping the momenta and integrating again returns to
    the start (to rounding); links stay unitary with unit determinant"""
    rng = np.random.default_rng(23)
    shape, nb, beta, eps, nlf = (4, 2, 3, 4), 2, 6.0, 0.07, 5
    full = (nb, 4, *shape, 3, 3)
    x = np.ascontiguousarray(osu3.random_su3(rng, full))
    v = np.ascontiguousarray(osu3.random_momentum(rng, full))
    dims = (ctypes.c_int * 4)(*shape)
    x1, v1, e1 = np.empty_like(x), np.empty_like(x), np.empty((nb, 4))
    emu.emu_hmc(ptr(x), ptr(v), ctypes.c_double(beta), ctypes.c_double(eps), ctypes.c_int(nlf), ptr(x1), ptr(v1), ptr(e1),
                ctypes.c_int(nb), dims)
    vm = np.ascontiguousarray(-v1)
    x2, v2, e2 = np.empty_like(x), np.empty_like(x), np.empty((nb, 4))
    emu.emu_hmc(ptr(x1), ptr(vm), ctypes.c_double(beta), ctypes.c_double(eps), ctypes.c_int(nlf), ptr(x2), ptr(v2), ptr(e2),
                ctypes.c_int(nb), dims)
    assert maxdiff(x2, x) < 1e-12 and maxdiff(-v2, v) < 1e-12
    assert np.allclose(e2[:, 2] + e2[:, 3], e1[:, 0] + e1[:, 1], rtol=1e-13)        # H returns to its start value
    # (the closed-form projectSU start is unitary to ~6e-13; exp(eps P) U must not make it worse)
    u0 = maxdiff(osu3.adj(x) @ x, np.broadcast_to(np.eye(3), x.shape))
    assert maxdiff(osu3.adj(x1) @ x1, np.broadcast_to(np.eye(3), x1.shape)) < u0 + 1e-13
    assert np.abs(osu3.det3(x1) - 1.0).max() < np.abs(osu3.det3(x) - 1.0).max() + 1e-13
    # second-order integrator: halving the step at fixed trajectory length divides the energy error by ~4
    dh = (e1[:, 2] + e1[:, 3]) - (e1[:, 0] + e1[:, 1])
    x3, v3, e3 = np.empty_like(x), np.empty_like(x), np.empty((nb, 4))
    emu.emu_hmc(ptr(x), ptr(v), ctypes.c_double(beta), ctypes.c_double(eps / 2), ctypes.c_int(2 * nlf), ptr(x3), ptr(v3),
                ptr(e3), ctypes.c_int(nb), dims)
    dh2 = (e3[:, 2] + e3[:, 3]) - (e3[:, 0] + e3[:, 1])
    ratio = np.abs(dh2) / np.abs(dh)
    assert np.all((ratio > 0.15) & (ratio < 0.35)), ratio


def test_improved_action_adjoints_match_torch_autograd(emu, golden_dir):
    """improved_action_adjoint_link (body of k_action_grad_c1) against torch autograd through a torch
    restatement of the reference's c1 action (plaquettes + `_rect_traces`, lattice.py:96-112,252-269) and
    through its force projectTAH(dsdx @ x^+) with dsdx detached (lattice.py:299-308)"""
    import torch
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    g = np.load(golden_dir / 'su3_c1_f64.npz')
    shape, nb, beta, c1 = [int(s) for s in g['shape']], g['x'].shape[0], float(g['beta']), float(g['c1'])
    lat = LatticeSU3(nb, shape, c1=c1)
    dims = (ctypes.c_int * 4)(*shape)

    def action(xt):      # S = -(beta/3) [(1 - 8 c1) sum Re tr P + c1 sum Re tr R]
        tr = lambda a: torch.diagonal(a, dim1=-2, dim2=-1).sum(-1)  # noqa: E731
        ps = 0.0
        for u in range(1, 4):
            for v in range(u):
                xu, xv = xt[:, u], xt[:, v]
                p = tr(xu @ xv.roll(-1, dims=u + 1) @ (xv @ xu.roll(-1, dims=v + 1)).mH)
                ps = ps + p.real.flatten(1).sum(1)
        return -(beta / 3.0) * (1 - 8 * c1) * ps + lat._rect_action(xt, beta)
    x = torch.from_numpy(g['x']).requires_grad_(True)
    s = action(x)
    assert np.allclose(s.detach().numpy(), g['action'], rtol=1e-12)
    rng = np.random.default_rng(4)
    gs = rng.standard_normal(nb)
    want, = torch.autograd.grad(s, x, grad_outputs=torch.from_numpy(gs))
    coef = np.ascontiguousarray(gs * (-beta / 3.0))
    xn = np.ascontiguousarray(g['x'])
    gx = np.empty_like(xn)
    emu.emu_action_grad_c1(ptr(xn), ptr(coef), ctypes.c_double(0.0), ctypes.c_double(c1), None, ptr(gx), ctypes.c_int(nb), dims)
    assert maxdiff(gx, want.numpy()) < 1e-12 * max(1.0, np.abs(want.numpy()).max())
    # force adjoint at fixed dsdx
    dsdx, = torch.autograd.grad(action(x).sum(), x)
    x2 = torch.from_numpy(g['x']).requires_grad_(True)
    y = dsdx.detach() @ x2.mH
    a = 0.5 * (y - y.mH)
    f = a - torch.diagonal(a, dim1=-2, dim2=-1).sum(-1)[..., None, None] / 3.0 * torch.eye(3, dtype=a.dtype)
    assert maxdiff(f.detach().numpy(), g['force']) < 1e-12
    gf = np.ascontiguousarray(rng.standard_normal(xn.shape) + 1j * rng.standard_normal(xn.shape))
    want2, = torch.autograd.grad(f, x2, grad_outputs=torch.from_numpy(gf))
    want2 = want2.resolve_conj()
    gx2 = np.empty_like(xn)
    emu.emu_action_grad_c1(ptr(xn), None, ctypes.c_double(-beta / 3.0), ctypes.c_double(c1), ptr(gf), ptr(gx2),
                           ctypes.c_int(nb), dims)
    assert maxdiff(gx2, want2.numpy()) < 1e-12 * max(1.0, np.abs(want2.numpy()).max())

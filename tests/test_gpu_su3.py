"""GPU tier (-m gpu): the CUDA SU(3) path, called through the C ABI
(l2hmc_b200.ops -> libl2b.so), against (i) the golden vectors frozen from the
reference's own PyTorch path and (ii) the numpy oracle on fresh seeded inputs,
plus size-independent properties at larger sizes.

Tolerances (north star: 1e-12 at fp64):
  * fields (links, momenta, force): 1e-12 absolute, entries are O(1);
  * per-chain sums (action, H): 1e-12 RELATIVE -- |S| ~ 6 beta V grows with the
    volume, so an absolute 1e-12 is below one ulp of the sum itself;
  * accept probability: 1e-12 * max(1, |H|) -- acc = exp(H0 - H1) inherits the
    absolute rounding of two sums of magnitude |H| (the reference differs from
    itself by as much between BLAS builds);
  * projectSU of an anti-Hermitian matrix (the reference feeds the FORCE through
    projectSU before su3_to_vec, dynamics.py:1154-1156) is ill conditioned: the
    oracle and the reference themselves differ by ~3e-10 there, so that one case
    is checked at 1e-7."""
import numpy as np
import pytest
import torch

from oracle import su3 as osu3, dynamics as od

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.detach().cpu().numpy()


def maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


@pytest.fixture(scope='module')
def ops():
    from l2hmc_b200 import ops as _ops
    return _ops


@pytest.fixture(scope='module')
def g(golden_dir):
    return np.load(golden_dir / 'su3_f64.npz')


def test_layout_roundtrip(ops, g):
    x = dev(g['x'])
    soa = ops.su3_aos_to_soa(x)
    back = ops.su3_soa_to_aos(soa)
    assert torch.equal(back, x)
    # planar layout: [b][mu][e][site]
    nb = x.shape[0]
    V = int(np.prod(g['shape']))
    want = g['x'].reshape(nb, 4, V, 9).transpose(0, 1, 3, 2)
    assert np.array_equal(host(soa).reshape(nb, 4, 9, V), want)


def test_wilson_loops_and_sums(ops, g):
    x = dev(g['x'])
    beta = float(g['beta'])
    V = int(np.prod(g['shape']))
    assert maxdiff(host(ops.su3_wilson_loops(x)), g['wloops']) < 1e-12
    s = host(ops.su3_plaq_sums(x))
    assert np.allclose(-beta / 3 * s[:, 0], g['action'], rtol=1e-12, atol=1e-12)
    assert maxdiff(s[:, 0] / (18 * V), g['plaqs']) < 1e-13
    assert maxdiff(s[:, 1] / (32 * np.pi ** 2), g['intQ']) < 1e-13
    assert maxdiff(s[:, 1] / (18 * V), g['sinQ']) < 1e-13


def test_force(ops, g):
    f, ps = ops.su3_force(dev(g['x']), float(g['beta']), want_plaq_sum=True)
    assert maxdiff(host(f), g['force']) < 1e-12
    assert np.allclose(-float(g['beta']) / 3 * host(ps), g['action'], rtol=1e-12, atol=1e-12)


def test_cold_start(ops, g):
    shape = [int(s) for s in g['shape']]
    V, beta = int(np.prod(shape)), float(g['beta'])
    cold = torch.eye(3, dtype=torch.complex128, device=DEV).expand(2, 4, *shape, 3, 3).contiguous()
    s = host(ops.su3_plaq_sums(cold))
    assert np.array_equal(s[:, 0], np.full(2, 18.0 * V)) and np.all(s[:, 1] == 0)
    assert np.allclose(-beta / 3 * s[:, 0], g['cold_action'], rtol=1e-15)
    assert float(ops.su3_force(cold, beta).abs().max()) == 0.0


def test_group_ops(ops, g):
    x, v, y = dev(g['x']), dev(g['v']), dev(g['y'])
    assert maxdiff(host(ops.su3_exp(v, 0.25)), g['expv']) < 1e-12
    assert maxdiff(host(ops.su3_exp(y)), g['expy']) < 1e-11 * np.abs(g['expy']).max()
    assert maxdiff(host(ops.su3_update_gauge(x, v, 0.1)), g['upd']) < 1e-12
    assert maxdiff(host(ops.su3_tah(y)), g['tah_y']) < 1e-15
    m, vec = ops.su3_project(y, want_matrix=True, want_vec=True)
    assert maxdiff(host(m), g['projsu_y']) < 1e-10
    assert maxdiff(host(vec), osu3.su3_to_vec(g['projsu_y'])) < 1e-9
    assert maxdiff(host(ops.su3_project(x, False, True)), g['vec_x']) < 1e-12
    assert maxdiff(host(ops.su3_project(dev(g['force']), False, True)), g['vec_f']) < 1e-7
    assert maxdiff(host(ops.su3_to_vec(dev(g['tah_y']))), osu3.su3_to_vec(g['tah_y'])) < 1e-14
    assert maxdiff(host(ops.su3_from_vec(dev(g['vec_x']))), g['vec2su3']) < 1e-15
    assert np.allclose(host(ops.su3_kinetic(v)), g['ke'], rtol=1e-12)
    a, mx = ops.su3_check(y)
    assert np.allclose(host(a), g['checksu_avg'], rtol=1e-12)
    assert np.allclose(host(mx), g['checksu_max'], rtol=1e-12)
    a, mx = ops.su3_check(x)
    assert float(mx.max()) < 1e-10


@pytest.mark.parametrize('key,xk,vk', [('hmc1', 'x', 'v'), ('hmc4', 'x', 'v'), ('hmcw', 'xw', 'vw')])
def test_hmc_trajectory_vs_reference_golden(ops, g, key, xk, vk):
    beta = float(g['beta'])
    nlf = int(g['hmcw_nlf']) if key == 'hmcw' else int(key[3:])
    xo, vo, en = ops.su3_hmc_trajectory(dev(g[xk]), dev(g[vk]), beta, float(g[f'{key}_eps']), nlf)
    assert maxdiff(host(xo), g[f'{key}_x']) < 1e-12
    assert maxdiff(host(vo), g[f'{key}_v']) < 1e-12
    en = host(en)
    h0, h1 = en[:, 0] + en[:, 1], en[:, 2] + en[:, 3]
    assert np.allclose(h0, g[f'{key}_h0'], rtol=1e-12, atol=1e-12)
    assert np.allclose(h1, g[f'{key}_h1'], rtol=1e-12, atol=1e-12)
    acc = np.exp(np.minimum(h0 - h1, 0.0))
    scale = max(1.0, float(np.abs(g[f'{key}_h0']).max()))
    assert maxdiff(acc, g[f'{key}_acc']) < 1e-12 * scale


@pytest.mark.parametrize('shape,nb,nlf', [([4, 4, 4, 4], 2, 1), ([4, 4, 4, 4], 3, 2), ([3, 5, 4, 7], 2, 4),
                                          ([8, 8, 8, 8], 5, 3)])
def test_trajectory_with_conversions_folded_in_is_bit_identical(ops, shape, nb, nlf):
    """default: the momenta are read from / written to the boundary layout by the trajectory's first / last
    launches and x_prop by its last drift (k_force_epx); `su3_fuse_conversions = 0`: four separate conversion
    kernels.  Same arithmetic in the same order: links and momenta agree bit for bit; the initial kinetic
    energy is summed over a different block partition (rounding of a sum only)."""
    from l2hmc_b200 import _lib
    rng = np.random.default_rng(7 * nb + nlf)
    full = (nb, 4, *shape, 3, 3)
    x, v = dev(osu3.random_su3(rng, full)), dev(osu3.random_momentum(rng, full))
    try:
        _lib.set_option('su3_fuse_conversions', 0)
        x0, v0, e0 = ops.su3_hmc_trajectory(x, v, 5.7, 0.07, nlf)
        _lib.set_option('su3_fuse_conversions', 1)
        x1, v1, e1 = ops.su3_hmc_trajectory(x, v, 5.7, 0.07, nlf)
    finally:
        _lib.set_option('su3_fuse_conversions', 1)
    assert torch.equal(x0, x1) and torch.equal(v0, v1)
    assert torch.equal(e0[:, 1:], e1[:, 1:])
    assert torch.allclose(e0[:, 0], e1[:, 0], rtol=1e-13, atol=1e-12)


@pytest.mark.parametrize('variant', [32, 33])
@pytest.mark.parametrize('shape,nb', [([4, 4, 4, 4], 3), ([8, 8, 8, 8], 2), ([3, 5, 4, 7], 2)])
def test_tma_staged_step_kernel_is_bit_identical(ops, shape, nb, variant):
    """force variants 32 / 33: the block's own momenta and links arrive through `cp.async.bulk.tensor.2d` + mbarrier
    (k_force_tma); same arithmetic as the default variant 22.  V % 32 != 0 ([3,5,4,7]) takes variant 22's kernels."""
    from l2hmc_b200 import _lib
    rng = np.random.default_rng(5 * nb + shape[0])
    full = (nb, 4, *shape, 3, 3)
    x, v = dev(osu3.random_su3(rng, full)), dev(osu3.random_momentum(rng, full))
    res = {}
    try:
        for var in (22, variant):
            _lib.set_option('su3_force_variant', var)
            for fuse in (1, 0):      # with the conversions folded into the end launches, and with all N_LF + 1 on the variant
                _lib.set_option('su3_fuse_conversions', fuse)
                res[(var, fuse)] = ops.su3_hmc_trajectory(x, v, 5.7, 0.07, 4)
    finally:
        _lib.set_option('su3_force_variant', 22)
        _lib.set_option('su3_fuse_conversions', 1)
    for fuse in (1, 0):
        for a, b in zip(res[(22, fuse)], res[(variant, fuse)]):
            assert torch.equal(a, b)


@pytest.mark.parametrize('shape,nb', [([2, 2, 2, 2], 3), ([4, 4, 4, 4], 2), ([6, 4, 2, 8], 1), ([3, 5, 4, 7], 2)])
def test_vs_oracle_on_fresh_inputs(ops, shape, nb):
    """ragged / odd / extent-2 lattices (extent 2: forward and backward neighbour coincide)"""
    rng = np.random.default_rng(sum(shape) + nb)
    full = (nb, 4, *shape, 3, 3)
    x = osu3.random_su3(rng, full)
    v = osu3.random_momentum(rng, full)
    beta = 5.5
    assert maxdiff(host(ops.su3_force(dev(x), beta)), osu3.grad_action(x, beta)) < 1e-12
    assert np.allclose(-beta / 3 * host(ops.su3_plaq_sums(dev(x)))[:, 0], osu3.action(x, beta), rtol=1e-12, atol=1e-11)
    s, acc = od.transition_kernel_hmc(od.SU3Ops, od.State(x, v, beta), 0.07, 3)
    xo, vo, en = ops.su3_hmc_trajectory(dev(x), dev(v), beta, 0.07, 3)
    assert maxdiff(host(xo), s.x) < 1e-12 and maxdiff(host(vo), s.v) < 1e-12
    en = host(en)
    assert np.allclose(en[:, 0], osu3.kinetic_energy(v), rtol=1e-12, atol=1e-11)
    assert np.allclose(en[:, 1], osu3.action(x, beta), rtol=1e-12, atol=1e-11)
    assert np.allclose(en[:, 2], osu3.kinetic_energy(s.v), rtol=1e-12, atol=1e-11)
    assert np.allclose(en[:, 3], osu3.action(s.x, beta), rtol=1e-12, atol=1e-11)


def test_masked_update_gauge_matches_l2hmc_x_update(ops, golden_dir):
    gl = np.load(golden_dir / 'su3_l2hmc_f64.npz')
    x, v = dev(gl['x']), dev(gl['v'])
    m = dev(gl['masks'][0])
    eps = float(gl['xeps'][0])
    eps = eps / (1.0 + eps)
    assert maxdiff(host(ops.su3_update_gauge(x, v, eps, mask=m)), gl['xfwd_x']) < 1e-12
    assert maxdiff(host(ops.su3_update_gauge(x, v, -eps, mask=m)), gl['xbwd_x']) < 1e-12
    # complement flag == passing 1 - m
    a = ops.su3_update_gauge(x, v, eps, mask=m, mask_complement=True)
    b = ops.su3_update_gauge(x, v, eps, mask=1.0 - m)
    assert torch.equal(a, b)


def test_vupdate_epilogue(ops):
    rng = np.random.default_rng(5)
    shape, nb = [2, 4, 2, 3], 2
    full = (nb, 4, *shape, 3, 3)
    v = osu3.random_momentum(rng, full)
    f = osu3.random_momentum(rng, full) * 2
    s, t, q = (rng.standard_normal((nb, 4 * 48 * 9)) * 0.3 for _ in range(3))
    eps = 0.1 / 1.1
    for sign in (+1, -1):
        out, ld = ops.su3_vupdate(dev(v), dev(f), dev(s), dev(t), dev(q), eps, sign)
        lj = sign * eps * s / 2.0
        es, eq = np.exp(lj).reshape(full), np.exp(eps * q).reshape(full)
        fn = f * eq + t.reshape(full)
        want = es * v - 0.5 * eps * fn if sign > 0 else es * (v + 0.5 * eps * fn)
        assert maxdiff(host(out), want) < 1e-13
        assert maxdiff(host(ld), lj.sum(1)) < 1e-12
    out, ld = ops.su3_vupdate(dev(v), dev(f), None, None, None, 0.2, +1)
    assert maxdiff(host(out), v - 0.1 * f) < 1e-14 and float(ld.abs().max()) == 0.0


def test_rand_momentum_distribution(ops):
    shape, nb = [4, 4, 4, 8], 4
    p, ke = ops.su3_rand_momentum(nb, shape, seed=1234, offset=0, device=DEV, want_ke=True)
    pn = host(p)
    assert maxdiff(pn, osu3.projectTAH(pn)) < 1e-15, 'momenta must be traceless anti-Hermitian'
    n2 = (np.abs(pn) ** 2).sum((-1, -2))
    assert abs(n2.mean() - 8.0) < 0.05, '<|P|_F^2> = 8 (group.py:125-126 convention)'
    vec = osu3.su3_to_vec(pn)
    assert abs(vec.mean()) < 0.01 and abs(vec.var() - 2.0) < 0.02, 'X^a = -2 tr[T^a X] are N(0, 2): tr T^aT^b = -1/2'
    assert np.allclose(host(ke), osu3.kinetic_energy(pn), rtol=1e-12, atol=1e-10)
    p2 = ops.su3_rand_momentum(nb, shape, seed=1234, offset=0, device=DEV)
    p3 = ops.su3_rand_momentum(nb, shape, seed=1234, offset=1, device=DEV)
    assert torch.equal(p, p2) and not torch.equal(p, p3)


def test_reversibility_and_unitarity_at_larger_size(ops):
    """size-independent properties at 8^4 x 8 chains: leapfrog is reversible and
    keeps links in SU(3) (2.2e-15 / checkSU <= 1e-12)"""
    shape, nb = [8, 8, 8, 8], 8
    torch.manual_seed(0)
    v = ops.su3_rand_momentum(nb, shape, seed=7, offset=0, device=DEV)
    cold = torch.eye(3, dtype=torch.complex128, device=DEV).expand(nb, 4, *shape, 3, 3).contiguous()
    x = ops.su3_update_gauge(cold, ops.su3_rand_momentum(nb, shape, seed=8, offset=0, device=DEV), 0.3)
    x1, v1, en = ops.su3_hmc_trajectory(x, v, 6.0, 0.05, 6)
    x2, v2, en2 = ops.su3_hmc_trajectory(x1, -v1, 6.0, 0.05, 6)
    assert float((x2 - x).abs().max()) < 1e-12
    assert float((v2 + v).abs().max()) < 1e-12
    assert float(ops.su3_check(x1)[1].max()) < 1e-12
    assert torch.allclose(en[:, 0] + en[:, 1], en2[:, 2] + en2[:, 3], rtol=1e-12)
    # energy conservation improves as eps^2
    dh = []
    for eps, n in ((0.1, 2), (0.05, 4), (0.025, 8)):
        _, _, e = ops.su3_hmc_trajectory(x, v, 6.0, eps, n)
        dh.append(float((e[:, 2] + e[:, 3] - e[:, 0] - e[:, 1]).abs().mean()))
    assert dh[1] < dh[0] / 3 and dh[2] < dh[1] / 3, dh


def test_accept_mix_is_bit_exact_select(ops, g):
    x0, x1 = dev(g['x']), dev(g['hmc4_x'])
    v0, v1 = dev(g['v']), dev(g['hmc4_v'])
    acc = torch.tensor([1.0, 0.0], device=DEV)
    xo, vo = ops.accept_mix(acc, [(x0, x1), (v0, v1)])
    st0, st1 = od.State(g['x'], g['v'], 0), od.State(g['hmc4_x'], g['hmc4_v'], 0)
    wx, wv, _ = od.accept_mix(np.array([0.9, 0.1]), np.array([0.5, 0.5]), st0, st1)
    assert np.array_equal(host(xo), wx) and np.array_equal(host(vo), wv)


@pytest.mark.parametrize('kernel', [False, True])
def test_rectangle_action_c1_matches_reference(golden_dir, kernel):
    """LatticeSU3(c1 != 0): loops, action, force and a plain-HMC trajectory against the reference's own
    c1 = -0.331 run.  kernel=False: plaquette part on the kernels + rectangle part as ATen ops;
    kernel=True: everything in the rectangle-staple kernel (l2b_su3_force_c1)"""
    from l2hmc_b200.configs import DynamicsConfig
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        g = np.load(golden_dir / 'su3_c1_f64.npz')
        shape, nb, beta, c1 = [int(s) for s in g['shape']], g['x'].shape[0], float(g['beta']), float(g['c1'])
        lat = LatticeSU3(nb, shape, c1=c1)
        if kernel and not lat.rect_kernel:
            pytest.skip('rectangle kernel disabled by L2B_RECT_KERNEL=0')
        lat.rect_kernel = kernel
        x, v = dev(g['x']), dev(g['v'])
        b = torch.tensor(beta)
        if kernel:
            from l2hmc_b200 import ops as _ops
            sums = _ops.su3_force_c1(x, beta, c1, want_force=False, want_sums=True)
            rs_want = g['rects'].real.reshape(12, nb, -1).sum(2).sum(0)
            assert np.all(np.abs(host(sums)[:, 1] - rs_want) <= 1e-12 * np.abs(rs_want).max())
            s2, f2 = lat.action_with_grad(x, b)
            assert np.all(np.abs(host(s2) - g['action']) <= 1e-12 * np.abs(g['action']))
            assert maxdiff(host(f2), g['force']) < 1e-12
        ps, rs = lat._wilson_loops(x, needs_rect=True)
        assert maxdiff(host(rs), g['rects']) < 1e-13
        assert np.all(np.abs(host(lat.action(x, b)) - g['action']) <= 1e-12 * np.abs(g['action']))
        assert maxdiff(host(lat.grad_action(x, b)), g['force']) < 1e-12
        cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=2, eps=0.05, eps_hmc=0.05,
                             verbose=False, use_split_xnets=False, use_separate_networks=False)
        dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
        with torch.no_grad():
            sp, met = dyn.transition_kernel_hmc(State(x, v, b), eps=0.05, nleapfrog=3)
        assert maxdiff(host(sp.x).reshape(g['hmc_x'].shape), g['hmc_x']) < 1e-12
        assert maxdiff(host(sp.v).reshape(g['hmc_v'].shape), g['hmc_v']) < 1e-12
        assert maxdiff(host(met['acc']), g['hmc_acc']) < 1e-10
    finally:
        torch.set_default_dtype(old)


def test_rectangle_kernel_gradients(golden_dir):
    """improved action under autograd on the adjoint kernel (l2b_su3_action_grad_c1) == the ATen-op path"""
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    g = np.load(golden_dir / 'su3_c1_f64.npz')
    shape, nb, beta, c1 = [int(s) for s in g['shape']], g['x'].shape[0], float(g['beta']), float(g['c1'])
    lat = LatticeSU3(nb, shape, c1=c1)
    gs = dev(np.array([0.7, -1.3]))
    gf = dev(np.random.default_rng(2).standard_normal(g['x'].shape) + 0j)
    res = {}
    for flag in (False, True):
        lat.rect_kernel_autograd = flag
        x = dev(g['x']).requires_grad_(True)
        s = lat.action(x, torch.tensor(beta))
        f = lat.grad_action(x, torch.tensor(beta))
        gx, = torch.autograd.grad([(s * gs).sum() + (f * gf.conj()).real.sum()], x)
        res[flag] = (host(s), host(f), host(gx))
    for a, b in zip(res[False], res[True]):
        assert maxdiff(a, b) < 1e-11 * max(1.0, float(np.abs(a).max()))

"""CPU tier: the numpy oracle (oracle/*.py) reproduces the golden vectors that
`oracle/make_golden.py` froze from the reference's own PyTorch path.  This is
what pins the oracle (the reference ships no tests)."""
import numpy as np
import pytest

from oracle import su3 as osu3, u1 as ou1, dynamics as od, network as onet


def maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))))


@pytest.fixture(scope='module')
def g(golden_dir):
    return np.load(golden_dir / 'su3_f64.npz')


def test_su3_observables(g):
    x, beta = g['x'], float(g['beta'])
    assert maxdiff(osu3.wilson_loops(x), g['wloops']) < 1e-13
    assert np.allclose(osu3.action(x, beta), g['action'], rtol=1e-12, atol=1e-12)
    assert maxdiff(osu3.plaqs(x), g['plaqs']) < 1e-14
    assert maxdiff(osu3.int_charges(x), g['intQ']) < 1e-14
    assert maxdiff(osu3.sin_charges(x), g['sinQ']) < 1e-14
    assert np.allclose(osu3.kinetic_energy(g['v']), g['ke'], rtol=1e-12)


def test_su3_force_is_reference_autograd_force(g):
    assert maxdiff(osu3.grad_action(g['x'], float(g['beta'])), g['force']) < 1e-13


def test_su3_cold_start_known_answers(g):
    shape = tuple(g['shape'])
    V = int(np.prod(shape))
    beta = float(g['beta'])
    cold = np.broadcast_to(np.eye(3, dtype=np.complex128), (2, 4, *shape, 3, 3)).copy()
    assert np.allclose(g['cold_action'], -6 * beta * V, rtol=1e-14)
    assert np.allclose(osu3.action(cold, beta), -6 * beta * V, rtol=1e-14)
    assert float(g['cold_force_max']) < 1e-14
    assert np.abs(osu3.grad_action(cold, beta)).max() < 1e-14
    assert np.allclose(osu3.plaqs(cold), 1.0)


def test_su3_group_ops(g):
    x, v, y = g['x'], g['v'], g['y']
    assert maxdiff(osu3.update_gauge(x, 0.1 * v), g['upd']) < 1e-14
    assert maxdiff(osu3.expm(0.25 * v), g['expv']) < 1e-14
    assert maxdiff(osu3.expm(y), g['expy']) < 1e-11 * np.abs(g['expy']).max()
    assert maxdiff(osu3.projectTAH(y), g['tah_y']) < 1e-15
    assert maxdiff(osu3.projectSU(y), g['projsu_y']) < 1e-10
    assert maxdiff(osu3.group_to_vec(x), g['vec_x']) < 1e-13
    # projectSU of an anti-Hermitian force is ill conditioned in the reference itself
    assert maxdiff(osu3.group_to_vec(g['force']), g['vec_f']) < 1e-7
    assert maxdiff(osu3.vec_to_su3(g['vec_x']), g['vec2su3']) < 1e-15
    a, m = osu3.checkSU(y)
    assert np.allclose(a, g['checksu_avg'], rtol=1e-13)
    assert np.allclose(m, g['checksu_max'], rtol=1e-13)


@pytest.mark.parametrize('key,xk,vk', [('hmc1', 'x', 'v'), ('hmc4', 'x', 'v'), ('hmcw', 'xw', 'vw')])
def test_su3_hmc_trajectory(g, key, xk, vk):
    beta = float(g['beta'])
    nlf = int(g['hmcw_nlf']) if key == 'hmcw' else int(key[3:])
    s, acc = od.transition_kernel_hmc(od.SU3Ops, od.State(g[xk], g[vk], beta), float(g[f'{key}_eps']), nlf)
    assert maxdiff(s.x, g[f'{key}_x']) < 1e-13
    assert maxdiff(s.v, g[f'{key}_v']) < 1e-13
    assert maxdiff(acc, g[f'{key}_acc']) < 1e-12
    if key == 'hmcw':
        assert np.all((g['hmcw_acc'] > 0) & (g['hmcw_acc'] < 1)), 'warm-start golden must exercise 0<acc<1'


def _spec_su3(gl):
    shape = [int(s) for s in gl['shape']]
    sd = {k[3:]: gl[k] for k in gl.files if k.startswith('sd/')}
    return od.L2HMCSpec(
        group='SU3', xshape=(2, 4, *shape, 3, 3), nleapfrog=int(gl['nlf']), xeps=list(gl['xeps']),
        veps=list(gl['veps']), masks=list(gl['masks']), state_dict=sd, activation='tanh',
        use_split_xnets=False, use_separate_networks=False, nw_x=tuple(gl['nw_x']), nw_v=tuple(gl['nw_v']))


def test_su3_l2hmc(golden_dir):
    gl = np.load(golden_dir / 'su3_l2hmc_f64.npz')
    spec = _spec_su3(gl)
    st = od.State(gl['x'], gl['v'], float(gl['beta']))
    s1, ld = od.update_v(spec, 0, st, True)
    assert maxdiff(s1.v, gl['vfwd_v']) < 1e-9 and maxdiff(ld, gl['vfwd_logdet']) < 1e-9
    s1, ld = od.update_v(spec, 1, st, False)
    assert maxdiff(s1.v, gl['vbwd_v']) < 1e-9 and maxdiff(ld, gl['vbwd_logdet']) < 1e-9
    s2, _ = od.update_x(spec, 0, st, spec.masks[0], True, True)
    assert maxdiff(s2.x, gl['xfwd_x']) < 1e-13
    s2, _ = od.update_x(spec, 0, st, spec.masks[0], True, False)
    assert maxdiff(s2.x, gl['xbwd_x']) < 1e-13
    sp, acc, sld = od.transition_kernel_fb(spec, st)
    assert maxdiff(sp.x, gl['fb_x']) < 1e-9
    assert maxdiff(sp.v, gl['fb_v']) < 1e-8
    assert maxdiff(acc, gl['fb_acc']) < 1e-8
    assert maxdiff(sld, gl['fb_sumlogdet']) < 1e-8


@pytest.mark.parametrize('tag,tol', [('f64', 1e-12), ('f32', 2e-5)])
def test_u1(golden_dir, tag, tol):
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    x, beta = gu['x'], float(gu['beta'])
    assert x.dtype == (np.float64 if tag == 'f64' else np.float32)
    assert maxdiff(ou1.wilson_loops(x), gu['wloops']) <= tol
    assert np.allclose(ou1.action(x, beta), gu['action'], rtol=tol)
    assert maxdiff(ou1.grad_action(x, beta), gu['force']) <= 10 * tol
    assert maxdiff(ou1.plaqs(x), gu['plaqs']) <= tol
    assert maxdiff(ou1.sin_charges(x), gu['sinQ']) <= tol
    assert maxdiff(ou1.int_charges(x), gu['intQ']) <= tol
    # "bit-exact integer topological charge": compare after rounding
    assert np.array_equal(np.round(ou1.int_charges(x)), np.round(gu['intQ']))
    assert maxdiff(ou1.wilson_loops4x4(x), gu['wloops4x4']) <= 10 * tol
    assert maxdiff(ou1.compat_proj(3 * x), gu['compat']) <= 10 * tol
    s, acc = od.transition_kernel_hmc(od.U1Ops, od.State(x, gu['v'], beta), 0.1, 5)
    assert maxdiff(s.x, gu['hmc_x']) <= 10 * tol and maxdiff(s.v, gu['hmc_v']) <= 10 * tol
    assert maxdiff(acc, gu['hmc_acc']) <= 100 * tol
    for name, conv in (('dense', None), ('conv', dict(filters=[4, 8, 8], sizes=[3, 2, 2], pool=[2, 2, 2]))):
        pre = f'{name}/'
        sd = {k[len(pre) + 3:]: gu[k] for k in gu.files if k.startswith(pre + 'sd/')}
        spec = od.L2HMCSpec(
            group='U1', xshape=(3, 2, *[int(s) for s in gu['shape']]), nleapfrog=2, xeps=list(gu[pre + 'xeps']),
            veps=list(gu[pre + 'veps']), masks=list(gu[pre + 'masks']), state_dict=sd, activation='leaky_relu',
            use_batch_norm=True, conv=conv)
        st = od.State(x, gu[pre + 'v'], beta)
        sp, acc, sld = od.transition_kernel_fb(spec, st)
        assert maxdiff(sp.x, gu[pre + 'fb_x']) <= 50 * tol
        assert maxdiff(sp.v, gu[pre + 'fb_v']) <= 50 * tol
        assert maxdiff(acc, gu[pre + 'fb_acc']) <= 50 * tol
        assert maxdiff(sld, gu[pre + 'fb_sumlogdet']) <= 50 * tol
        s2, ld2 = od.update_x(spec, 0, st, spec.masks[0], True, True)
        assert maxdiff(s2.x, gu[pre + 'xfwd_x']) <= 10 * tol and maxdiff(ld2, gu[pre + 'xfwd_logdet']) <= 10 * tol
        s3, ld3 = od.update_x(spec, 1, st, spec.masks[0], False, False)
        assert maxdiff(s3.x, gu[pre + 'xbwd_x']) <= 10 * tol and maxdiff(ld3, gu[pre + 'xbwd_logdet']) <= 10 * tol
        s4, ld4 = od.update_v(spec, 0, st, True)
        assert maxdiff(s4.v, gu[pre + 'vfwd_v']) <= 10 * tol and maxdiff(ld4, gu[pre + 'vfwd_logdet']) <= 10 * tol


def test_su3_rectangle_action_oracle_matches_reference(golden_dir):
    """c1 != 0 (DBW2 rectangle term, lattice.py:96-112,180-196,252-269)"""
    g = np.load(golden_dir / 'su3_c1_f64.npz')
    assert maxdiff(osu3.rect_traces(g['x']), g['rects']) < 1e-13
    got = osu3.action(g['x'], float(g['beta']), float(g['c1']))
    assert np.all(np.abs(got - g['action']) <= 1e-12 * np.abs(g['action']))


def test_su3_hmc_with_improved_action_uses_the_wilson_force(golden_dir):
    """Reference semantics pinned: `Dynamics` integrates with the force of ITS OWN lattice, built with the
    default c1 = 0 (dynamics.py:134,1499), while the energies of the accept step come from `potential_fn`
    (here the c1 = -0.331 action, dynamics.py:1489-1491).  The oracle's plain-Wilson trajectory reproduces the
    reference's proposal, and the accept probability follows from the improved action."""
    g = np.load(golden_dir / 'su3_c1_f64.npz')
    beta, c1 = float(g['beta']), float(g['c1'])
    x0, v0 = g['hmc2_x0'], g['hmc2_v0']
    s, _ = od.transition_kernel_hmc(od.SU3Ops, od.State(x0, v0, beta), 0.01, 3)
    assert maxdiff(s.x.reshape(x0.shape), g['hmc2_x'].reshape(x0.shape)) < 1e-12
    assert maxdiff(s.v.reshape(x0.shape), g['hmc2_v'].reshape(x0.shape)) < 1e-12
    h0 = osu3.kinetic_energy(v0) + osu3.action(x0, beta, c1)
    h1 = osu3.kinetic_energy(s.v.reshape(x0.shape)) + osu3.action(s.x.reshape(x0.shape), beta, c1)
    assert np.allclose(h0, g['hmc2_h0'], rtol=1e-12) and np.allclose(h1, g['hmc2_h1'], rtol=1e-12)
    acc = np.exp(np.minimum(h0 - h1, 0.0))
    assert np.allclose(acc, g['hmc2_acc'], rtol=1e-9) and 0 < acc.min() < 0.01 < acc.max() < 1
    # with the Wilson energies the accept probability would be a different number
    hw0 = osu3.kinetic_energy(v0) + osu3.action(x0, beta)
    hw1 = osu3.kinetic_energy(s.v.reshape(x0.shape)) + osu3.action(s.x.reshape(x0.shape), beta)
    assert not np.allclose(np.exp(np.minimum(hw0 - hw1, 0.0)), g['hmc2_acc'], rtol=1e-3)


@pytest.mark.parametrize('tag,tol', [('f64', 1e-12), ('f32', 2e-5)])
def test_u1_unmerged_kernel_and_verbose_histories(golden_dir, tag, tol):
    """transition_kernel (merge_directions = False, swapped states in the accept probability) and the
    verbose per-step metrics of all three kernels (dynamics.py:865-898,915-1063) against the reference"""
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    pre = 'dense/'
    sd = {k[len(pre) + 3:]: gu[k] for k in gu.files if k.startswith(pre + 'sd/')}
    spec = od.L2HMCSpec(group='U1', xshape=(3, 2, *[int(s) for s in gu['shape']]), nleapfrog=2,
                        xeps=list(gu[pre + 'xeps']), veps=list(gu[pre + 'veps']), masks=list(gu[pre + 'masks']),
                        state_dict=sd, activation='leaky_relu', use_batch_norm=True)
    st = od.State(gu['x'], gu[pre + 'v'], float(gu['beta']))
    for key, fwd in (('tkf', True), ('tkb', False)):
        sp, acc, sld = od.transition_kernel(spec, st, fwd)
        assert maxdiff(sp.x.reshape(gu[f'{pre}{key}_x'].shape), gu[f'{pre}{key}_x']) <= 50 * tol
        assert maxdiff(sp.v, gu[f'{pre}{key}_v']) <= 50 * tol
        assert maxdiff(sld, gu[f'{pre}{key}_sumlogdet']) <= 50 * tol
        assert maxdiff(acc, gu[f'{pre}{key}_acc']) <= 200 * tol
    # swapped: the un-merged acc is exp(min(H_final - H_init + sld, 0)), not the usual H_init - H_final
    usual = od.accept_prob(spec.g, st, od.State(sp.x.reshape(st.x.shape), sp.v, st.beta), sld)
    assert maxdiff(usual, gu[pre + 'tkb_acc']) > 1e-3
    htol = 500 * tol * max(1.0, float(np.abs(gu[pre + 'vfb/energy']).max()))
    for key, (sp, h) in (('vfb', od.transition_kernel_fb_verbose(spec, st)),
                         ('vtk', od.transition_kernel_verbose(spec, st, True)),
                         ('vhmc', od.transition_kernel_hmc_verbose(spec, st, 0.1, 3))):
        want = {k[len(pre) + len(key) + 1:]: gu[k] for k in gu.files if k.startswith(f'{pre}{key}/')}
        assert set(want) == set(h), (key, sorted(set(want) ^ set(h)))
        for k, w in want.items():
            assert np.shape(h[k]) == w.shape, (key, k, np.shape(h[k]), w.shape)
            assert maxdiff(h[k], w) <= htol, (key, k)
        assert maxdiff(sp.x.reshape(gu[f'{pre}{key}_x'].shape), gu[f'{pre}{key}_x']) <= 50 * tol

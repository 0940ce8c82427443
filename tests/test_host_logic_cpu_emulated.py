"""CPU tier: the HOST logic of the `Dynamics` mirror, end to end, with the U(1) kernels replaced by CPU torch
stand-ins (tests/cpu_emulation.py, test infrastructure).  Same assertions as the GPU tier's golden tests: the
merged sweep, the un-merged kernel (swapped accept states), the verbose histories of all three kernels, plain
HMC, and the public `forward` contract with and without `merge_directions`."""
import numpy as np
import pytest
import torch

from tests.cpu_emulation import u1_host_logic_on_cpu


def maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


def host(t):
    return t.detach().numpy()


def _dynamics(gu, name, verbose, merge=True):
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, ConvolutionConfig, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    shape, nb, nlf = [int(s) for s in gu['shape']], 3, 2
    pre = f'{name}/'
    cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=nlf, eps=0.1, eps_hmc=0.1, use_ncp=True,
                         verbose=verbose, use_split_xnets=True, use_separate_networks=True, merge_directions=merge)
    conv = ConvolutionConfig(filters=[4, 8, 8], sizes=[3, 2, 2], pool=[2, 2, 2]) if name == 'conv' else None
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=[16, 12], activation_fn='leaky_relu', dropout_prob=0.2,
                                                      use_batch_norm=True),
                         conv_config=conv, net_weights=None)
    lat = LatticeU1(nb, shape)
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    sd = {k[len(pre) + 3:]: torch.from_numpy(gu[k]) for k in gu.files if k.startswith(pre + 'sd/')}
    res = dyn.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys
    dyn.masks = [torch.from_numpy(m) for m in gu[pre + 'masks']]
    dyn.eval()
    return dyn


@pytest.fixture()
def emulated(monkeypatch):
    with u1_host_logic_on_cpu(monkeypatch):
        yield


@pytest.fixture()
def dtype_of():
    old = torch.get_default_dtype()
    yield lambda tag: torch.set_default_dtype(torch.float64 if tag == 'f64' else torch.float32)
    torch.set_default_dtype(old)


@pytest.mark.parametrize('tag,tol', [('f64', 1e-11), ('f32', 2e-5)])
@pytest.mark.parametrize('name', ['dense', 'conv'])
def test_merged_sweep_and_single_updates(golden_dir, emulated, dtype_of, tag, tol, name):
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    dtype_of(tag)
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    pre = f'{name}/'
    dyn = _dynamics(gu, name, verbose=False)
    st = State(torch.from_numpy(gu['x']), torch.from_numpy(gu[pre + 'v']), torch.tensor(float(gu['beta'])))
    with torch.no_grad():
        m, _ = dyn._get_mask(0)
        s2, ld2 = dyn._update_x_fwd(0, st, m, first=True)
        assert maxdiff(host(s2.x), gu[pre + 'xfwd_x']) <= 10 * tol and maxdiff(host(ld2), gu[pre + 'xfwd_logdet']) <= 10 * tol
        s3, ld3 = dyn._update_x_bwd(1, st, m, first=False)
        assert maxdiff(host(s3.x), gu[pre + 'xbwd_x']) <= 10 * tol and maxdiff(host(ld3), gu[pre + 'xbwd_logdet']) <= 10 * tol
        s4, ld4 = dyn._update_v_fwd(0, st)
        assert maxdiff(host(s4.v), gu[pre + 'vfwd_v']) <= 10 * tol and maxdiff(host(ld4), gu[pre + 'vfwd_logdet']) <= 10 * tol
        sp, met = dyn.transition_kernel_fb(st)
    assert maxdiff(host(sp.x), gu[pre + 'fb_x']) <= 50 * tol
    assert maxdiff(host(sp.v).reshape(3, -1), gu[pre + 'fb_v']) <= 50 * tol
    assert maxdiff(host(met['acc']), gu[pre + 'fb_acc']) <= 50 * tol
    assert maxdiff(host(met['sumlogdet']), gu[pre + 'fb_sumlogdet']) <= 50 * tol


@pytest.mark.parametrize('tag,tol', [('f64', 1e-11), ('f32', 2e-5)])
def test_unmerged_kernel_and_plain_hmc(golden_dir, emulated, dtype_of, tag, tol):
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    dtype_of(tag)
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    pre = 'dense/'
    dyn = _dynamics(gu, 'dense', verbose=False)
    st = State(torch.from_numpy(gu['x']), torch.from_numpy(gu[pre + 'v']), torch.tensor(float(gu['beta'])))
    for key, fwd in (('tkf', True), ('tkb', False)):
        with torch.no_grad():
            sp, met = dyn.transition_kernel(st, forward=fwd)
        assert maxdiff(host(sp.x).reshape(gu[f'{pre}{key}_x'].shape), gu[f'{pre}{key}_x']) <= 50 * tol
        assert maxdiff(host(sp.v).reshape(3, -1), gu[f'{pre}{key}_v']) <= 50 * tol
        assert maxdiff(host(met['sumlogdet']), gu[f'{pre}{key}_sumlogdet']) <= 50 * tol
        assert maxdiff(host(met['acc']), gu[f'{pre}{key}_acc']) <= 200 * tol
    with torch.no_grad():
        sp, met = dyn.transition_kernel_hmc(State(st.x, torch.from_numpy(gu['v']), st.beta), eps=0.1, nleapfrog=5)
    assert maxdiff(host(sp.x).reshape(3, -1), gu['hmc_x']) <= 10 * tol and maxdiff(host(sp.v).reshape(3, -1), gu['hmc_v']) <= 10 * tol
    assert maxdiff(host(met['acc']), gu['hmc_acc']) <= 200 * tol * max(1.0, float(np.abs(gu['hmc_h0']).max()))


@pytest.mark.parametrize('tag,tol', [('f64', 1e-11), ('f32', 2e-5)])
def test_verbose_histories(golden_dir, emulated, dtype_of, tag, tol):
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    dtype_of(tag)
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    pre = 'dense/'
    dyn = _dynamics(gu, 'dense', verbose=True)
    st = State(torch.from_numpy(gu['x']), torch.from_numpy(gu[pre + 'v']), torch.tensor(float(gu['beta'])))
    htol = 500 * tol * max(1.0, float(np.abs(gu[pre + 'vfb/energy']).max()))
    with torch.no_grad():
        runs = (('vfb', dyn.transition_kernel_fb(st)), ('vtk', dyn.transition_kernel(st, forward=True)),
                ('vhmc', dyn.transition_kernel_hmc(st, eps=0.1, nleapfrog=3)))
    for key, (sp, h) in runs:
        want = {k[len(pre) + len(key) + 1:]: gu[k] for k in gu.files if k.startswith(f'{pre}{key}/')}
        got = {k: v for k, v in h.items() if isinstance(v, torch.Tensor)}
        assert set(want) == set(got), (key, sorted(set(want) ^ set(got)))
        for k, w in want.items():
            assert tuple(got[k].shape) == w.shape, (key, k, tuple(got[k].shape), w.shape)
            assert maxdiff(host(got[k]), w) <= htol, (key, k)
        assert maxdiff(host(sp.x).reshape(gu[f'{pre}{key}_x'].shape), gu[f'{pre}{key}_x']) <= 50 * tol


@pytest.mark.parametrize('merge', [True, False])
def test_forward_contract(golden_dir, emulated, dtype_of, merge):
    dtype_of('f64')
    gu = np.load(golden_dir / 'u1_f64.npz')
    dyn = _dynamics(gu, 'dense', verbose=False, merge=merge)
    x, beta = torch.from_numpy(gu['x']), torch.tensor(float(gu['beta']))
    torch.manual_seed(5)
    with torch.no_grad():
        xout, met = dyn((x, beta))
    nb = x.shape[0]
    assert tuple(xout.shape) == (nb, dyn.xdim)
    assert met['acc_mask'].dtype == torch.float32 and tuple(met['acc'].shape) == (nb,)
    mc = met['mc_states']
    xo, xp, xi = host(xout), host(mc.proposed.x).reshape(nb, -1), host(mc.init.x).reshape(nb, -1)
    ma = host(met['acc_mask'])
    for b in range(nb):
        assert np.array_equal(xo[b], xp[b] if ma[b] == 1.0 else xi[b])
    assert np.all((host(met['sumlogdet']) == 0) | (ma == 1.0))
    # apply_transition_both: forward and backward proposals mixed per chain
    with torch.no_grad():
        xo2, met2 = dyn.apply_transition_both((x, beta))
    assert tuple(xo2.shape) == (nb, dyn.xdim) and met2['acc_mask'].dtype == torch.float32
    xo_hmc, met_h = dyn.apply_transition_hmc((x, beta), eps=0.1, nleapfrog=3)
    assert tuple(xo_hmc.shape) == (nb, dyn.xdim) and 'sumlogdet' in met_h


def test_trainer_eval_and_hmc_steps(golden_dir, emulated, dtype_of):
    """Trainer.hmc_step / eval_step (trainers/pytorch/trainer.py:904-956): `compat_proj` at the top of the step
    (appendix B trap 9), the loss on (x_init, proposed x, acc), x_out detached and flattened"""
    from l2hmc_b200.configs import LossConfig
    from l2hmc_b200.loss.pytorch.loss import LatticeLoss
    from l2hmc_b200.trainers.pytorch.trainer import Trainer
    dtype_of('f64')
    gu = np.load(golden_dir / 'u1_f64.npz')
    dyn = _dynamics(gu, 'dense', verbose=False)
    lcfg = LossConfig(use_mixed_loss=True, charge_weight=0.01)
    tr = Trainer(dyn, loss_config=lcfg)
    x = torch.from_numpy(gu['x']) * 3.0                  # outside [-pi, pi): the step must wrap it first
    beta = torch.tensor(float(gu['beta']))
    for wraps, step in ((True, tr.eval_step), (False, lambda inp: tr.hmc_step(inp, eps=0.1, nleapfrog=3))):
        torch.manual_seed(11)
        xo, met = step((x, beta))
        assert tuple(xo.shape) == (3, dyn.xdim) and not xo.requires_grad
        # the L2HMC x-update wraps its output; plain HMC drifts x + eps v without wrapping, like the reference
        assert (float(xo.abs().max()) <= np.pi) or not wraps
        assert float(xo.abs().max()) < np.pi + 3.0           # but it started from the WRAPPED input (|3 x| reaches 9)
        assert 'mc_states' not in met and met['acc_mask'].dtype == torch.float32
        assert torch.isfinite(met['loss']) and met['loss'].dim() == 0
    # the loss is LatticeLoss of the wrapped input, the proposal and acc
    torch.manual_seed(11)
    xi = dyn.g.compat_proj(x)
    xo2, m2 = dyn((xi, beta))
    want = LatticeLoss(dyn.lattice, lcfg)(x_init=xi, x_prop=m2['mc_states'].proposed.x, acc=m2['acc'])
    torch.manual_seed(11)
    _, met = tr.eval_step((x, beta))
    assert float((met['loss'] - want).abs()) < 1e-12


@pytest.fixture()
def emulated_su3(monkeypatch):
    from tests.cpu_emulation import su3_host_logic_on_cpu
    with su3_host_logic_on_cpu(monkeypatch):
        yield


def test_su3_l2hmc_sweep_host_logic(golden_dir, emulated_su3, dtype_of):
    """the boundary-layout SU(3) sweep (vnet on su3_to_vec(projectSU(.)) of x and of the force, masked exp(eps v)
    x-update, logdet from s only) against the reference golden, kernels replaced by oracle-backed stand-ins"""
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, NetWeights, NetWeight, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    dtype_of('f64')
    gl = np.load(golden_dir / 'su3_l2hmc_f64.npz')
    shape, nb, nlf = [int(s) for s in gl['shape']], 2, int(gl['nlf'])
    cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=nlf, eps=0.05, eps_hmc=0.1,
                         verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=[8], activation_fn='tanh', dropout_prob=0.0, use_batch_norm=False),
                         conv_config=None, net_weights=NetWeights(x=NetWeight(0., 1., 1.), v=NetWeight(1., 1., 1.)))
    dyn = Dynamics(potential_fn=LatticeSU3(nb, shape).action, config=cfg, network_factory=fac)
    sd = {k[3:]: torch.from_numpy(gl[k]) for k in gl.files if k.startswith('sd/')}
    res = dyn.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys
    dyn.masks = [torch.from_numpy(m) for m in gl['masks']]
    dyn.eval()
    st = State(torch.from_numpy(gl['x']), torch.from_numpy(gl['v']), torch.tensor(float(gl['beta'])))
    with torch.no_grad():
        s1, ld = dyn._update_v_fwd(0, st)
        assert maxdiff(np.abs(host(s1.v) - gl['vfwd_v']), 0) < 1e-8 and maxdiff(host(ld), gl['vfwd_logdet']) < 1e-8
        m, _ = dyn._get_mask(0)
        s2, ld2 = dyn._update_x_fwd(0, st, m, first=True)
        assert np.abs(host(s2.x) - gl['xfwd_x']).max() < 1e-12 and float(ld2.abs().max()) == 0
        sp, met = dyn.transition_kernel_fb(st)
    assert np.abs(host(sp.x) - gl['fb_x']).max() < 1e-8 and np.abs(host(sp.v) - gl['fb_v']).max() < 1e-8
    assert maxdiff(host(met['acc']), gl['fb_acc']) < 1e-8 and maxdiff(host(met['sumlogdet']), gl['fb_sumlogdet']) < 1e-8


@pytest.mark.parametrize('kernel', [True, False])
def test_su3_hmc_with_improved_action_host_logic(golden_dir, emulated_su3, dtype_of, kernel):
    """plain HMC with potential_fn = improved action: trajectory from the Wilson force, acceptance from
    potential_fn (the `_potential_is_wilson` branch), both with the rectangle kernel routed in and with ATen ops"""
    from l2hmc_b200.configs import DynamicsConfig
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    dtype_of('f64')
    g = np.load(golden_dir / 'su3_c1_f64.npz')
    shape, nb, beta, c1 = [int(s) for s in g['shape']], g['x'].shape[0], float(g['beta']), float(g['c1'])
    lat = LatticeSU3(nb, shape, c1=c1)
    lat.rect_kernel = kernel
    cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=2, eps=0.05, eps_hmc=0.05,
                         verbose=False, use_split_xnets=False, use_separate_networks=False)
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
    assert not dyn._potential_is_wilson()
    st = State(torch.from_numpy(g['hmc2_x0']), torch.from_numpy(g['hmc2_v0']), torch.tensor(beta))
    with torch.no_grad():
        sp, met = dyn.transition_kernel_hmc(st, eps=0.01, nleapfrog=3)
        h0 = dyn.hamiltonian(st)
    assert np.abs(host(sp.x).reshape(g['hmc2_x0'].shape) - g['hmc2_x'].reshape(g['hmc2_x0'].shape)).max() < 1e-12
    assert np.abs(host(sp.v).reshape(g['hmc2_x0'].shape) - g['hmc2_v'].reshape(g['hmc2_x0'].shape)).max() < 1e-12
    assert np.allclose(host(h0), g['hmc2_h0'], rtol=1e-12)
    assert np.allclose(host(met['acc']), g['hmc2_acc'], rtol=1e-7)
    # and the plain Wilson potential takes the kernel's own energies
    dyn_w = Dynamics(potential_fn=LatticeSU3(nb, shape).action, config=cfg, network_factory=None)
    assert dyn_w._potential_is_wilson()
    with torch.no_grad():
        sp_w, met_w = dyn_w.transition_kernel_hmc(st, eps=0.01, nleapfrog=3)
    assert torch.equal(sp_w.x, sp.x) and not np.allclose(host(met_w['acc']), g['hmc2_acc'], rtol=1e-3)
    # lattice-level c1 API through the same routing
    x = torch.from_numpy(g['x'])
    with torch.no_grad():
        assert np.allclose(host(lat.action(x, torch.tensor(beta))), g['action'], rtol=1e-12)
        assert np.abs(host(lat.grad_action(x, torch.tensor(beta))) - g['force']).max() < 1e-12
        s2, f2 = lat.action_with_grad(x, torch.tensor(beta))
        assert np.allclose(host(s2), g['action'], rtol=1e-12) and np.abs(host(f2) - g['force']).max() < 1e-12


def test_su3_force_reuse_gives_identical_sweep_with_fewer_force_evaluations(golden_dir, emulated_su3, dtype_of):
    """`reuse_force = 'always'` on the boundary-layout sweep: same proposal, acceptance and log-Jacobian bit for
    bit, 2 nlf + 1 instead of 4 nlf force evaluations and projections per sweep"""
    from l2hmc_b200 import ops
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, NetWeights, NetWeight, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    dtype_of('f64')
    gl = np.load(golden_dir / 'su3_l2hmc_f64.npz')
    shape, nb, nlf = [int(s) for s in gl['shape']], 2, int(gl['nlf'])
    cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=nlf, eps=0.05, eps_hmc=0.1,
                         verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=[8], activation_fn='tanh', dropout_prob=0.0, use_batch_norm=False),
                         conv_config=None, net_weights=NetWeights(x=NetWeight(0., 1., 1.), v=NetWeight(1., 1., 1.)))
    dyn = Dynamics(potential_fn=LatticeSU3(nb, shape).action, config=cfg, network_factory=fac)
    dyn.load_state_dict({k[3:]: torch.from_numpy(gl[k]) for k in gl.files if k.startswith('sd/')}, strict=False)
    dyn.masks = [torch.from_numpy(m) for m in gl['masks']]
    dyn.eval()
    st = State(torch.from_numpy(gl['x']), torch.from_numpy(gl['v']), torch.tensor(float(gl['beta'])))
    counts = {'force': 0, 'vec': 0}
    force, vec = ops.su3_force, ops.su3_project_vec

    def counting_force(*a, **k):
        counts['force'] += 1
        return force(*a, **k)

    def counting_vec(*a, **k):
        counts['vec'] += 1
        return vec(*a, **k)
    ops.su3_force, ops.su3_project_vec = counting_force, counting_vec       # (monkeypatch restores the originals)
    res = {}
    for mode in ('never', 'always'):
        dyn.reuse_force = mode
        counts.update(force=0, vec=0)
        with torch.no_grad():
            sp, met = dyn.transition_kernel_fb(st)
        res[mode] = (sp.x, sp.v, met['acc'], met['sumlogdet'], dict(counts))
    assert res['never'][4] == {'force': 4 * nlf, 'vec': 8 * nlf}
    assert res['always'][4] == {'force': 2 * nlf + 1, 'vec': 2 * (2 * nlf + 1)}
    for a, b in zip(res['never'][:4], res['always'][:4]):
        assert torch.equal(a, b)
    assert np.abs(host(res['always'][0]) - gl['fb_x']).max() < 1e-8


def test_the_late_gpu_tests_themselves_run_clean_on_the_emulation(golden_dir, monkeypatch):
    """The GPU test files added after the round's last GPU run (tests/test_gpu_zz_*.py) are executed here, body for
    body, on the CPU stand-ins (their DEV constant pointed at the CPU), so that a slip in the TEST code cannot be
    what fails on the GPU box"""
    from tests.cpu_emulation import su3_host_logic_on_cpu
    import tests.test_gpu_zz_unmerged_and_verbose as zu
    import tests.test_gpu_zz_improved_action_acc as zi
    monkeypatch.setattr(zu, 'DEV', 'cpu')
    monkeypatch.setattr(zi, 'DEV', 'cpu')
    old = torch.get_default_dtype()
    try:
        with u1_host_logic_on_cpu(monkeypatch):
            for tag, tol in (('f64', 1e-11), ('f32', 2e-5)):
                zu.test_unmerged_transition_kernel_matches_reference(golden_dir, tag, tol)
                zu.test_verbose_histories_match_reference(golden_dir, tag, tol)
            zu.test_forward_without_merge_directions_goes_through_apply_transition(golden_dir)
        with su3_host_logic_on_cpu(monkeypatch):
            for kernel in (True, False):
                zi.test_hmc_accepts_with_the_energies_of_potential_fn(golden_dir, kernel)
            import tests.test_gpu_su3 as zs                  # improved action under autograd
            monkeypatch.setattr(zs, 'DEV', 'cpu')
            zs.test_rectangle_kernel_gradients(golden_dir)
            import tests.test_gpu_reuse_force as zr          # its no-autocast body runs here
            monkeypatch.setattr(zr, 'DEV', 'cpu')
            zr.test_su3_fb_sweep_is_bit_identical_and_evaluates_fewer_forces(golden_dir, False)
            zr.test_su3_gradients_with_reuse_equal_default(golden_dir)
    finally:
        torch.set_default_dtype(old)


@pytest.mark.parametrize('tag,rtol', [('f64', 1e-9), ('f32', 5e-3)])
@pytest.mark.parametrize('name', ['dense', 'conv'])
def test_u1_training_gradients_through_the_autograd_wiring(golden_dir, monkeypatch, tag, rtol, name):
    """The body of the GPU tier's gradient test (tests/test_gpu_training.py) on the CPU stand-ins (adjoints = torch
    vjps of the stand-ins): every parameter / step-size / input gradient our `torch.autograd.Function` wiring
    produces must equal the reference's autograd goldens"""
    import tests.test_gpu_training as tg
    monkeypatch.setattr(tg, 'DEV', 'cpu')
    old = torch.get_default_dtype()
    try:
        with u1_host_logic_on_cpu(monkeypatch):
            tg.test_u1_l2hmc_gradients_match_reference_autograd(golden_dir, torch.set_default_dtype, tag, rtol, name)
    finally:
        torch.set_default_dtype(old)


def test_gpu_tier_bodies_that_only_need_host_logic_run_on_the_emulation(golden_dir, monkeypatch):
    """Regression net for work done without a GPU: the bodies of the GPU tier's Dynamics / Trainer tests whose
    kernels have stand-ins are executed here on the CPU (same assertions, same goldens)"""
    from tests.cpu_emulation import su3_host_logic_on_cpu
    import tests.test_gpu_dynamics as td
    import tests.test_gpu_trainer as tt
    for mod in (td, tt):
        monkeypatch.setattr(mod, 'DEV', 'cpu')
    old = torch.get_default_dtype()
    try:
        with u1_host_logic_on_cpu(monkeypatch):
            for tag, tol in (('f64', 1e-11), ('f32', 2e-5)):
                for name in ('dense', 'conv'):
                    td.test_u1_l2hmc_matches_reference(golden_dir, torch.set_default_dtype, tag, tol, name, 'never')
            tt.test_u1_trainer_steps(torch.set_default_dtype)
        with su3_host_logic_on_cpu(monkeypatch):
            td.test_su3_l2hmc_matches_reference(golden_dir, torch.set_default_dtype)
            tt.test_su3_trainer_steps(torch.set_default_dtype)
    finally:
        torch.set_default_dtype(old)


def test_u1_gradients_are_unchanged_by_force_reuse(golden_dir, emulated, dtype_of):
    """training with `reuse_force = 'always'`: the shared force tensor receives the cotangents of both v-updates
    it feeds, so every gradient equals the default's (to rounding: the sums are formed in a different order)"""
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    dtype_of('f64')
    gu = np.load(golden_dir / 'u1_f64.npz')
    pre = 'dense/'
    dyn = _dynamics(gu, 'dense', verbose=False)
    n = {'force': 0}
    orig = dyn.grad_potential

    def counting(x, beta):
        n['force'] += 1
        return orig(x, beta)
    dyn.grad_potential = counting
    res = {}
    for mode in ('never', 'always'):
        dyn.reuse_force = mode
        dyn.zero_grad(set_to_none=True)
        n['force'] = 0
        x = torch.from_numpy(gu['x']).requires_grad_(True)
        st = State(x, torch.from_numpy(gu[pre + 'v']), torch.tensor(float(gu['beta'])))
        sp, met = dyn.transition_kernel_fb(st)
        loss = (met['acc'] * sp.x.flatten(1).cos().sum(1)).sum() + met['sumlogdet'].sum()
        loss.backward()
        res[mode] = (float(loss), x.grad.clone(), {k: p.grad.clone() for k, p in dyn.named_parameters() if p.grad is not None},
                     n['force'])
    assert res['never'][3] == 8 and res['always'][3] == 5            # 4 nlf vs 2 nlf + 1 at nlf = 2
    assert res['never'][0] == res['always'][0]
    assert float((res['never'][1] - res['always'][1]).abs().max()) <= 1e-12 * float(res['never'][1].abs().max())
    assert set(res['never'][2]) == set(res['always'][2]) and len(res['never'][2]) > 40
    for k, g0 in res['never'][2].items():
        assert float((g0 - res['always'][2][k]).abs().max()) <= 1e-12 * max(1e-30, float(g0.abs().max())), k


def test_su3_training_gradients_through_the_autograd_wiring(golden_dir, monkeypatch):
    """The SU(3) gradient-golden test of the GPU tier on the CPU: adjoint stand-ins are torch vjps of smooth
    restatements (polar factor by Newton iteration, torch.matrix_exp), i.e. independent of the kernels' closed
    forms; what is exercised is our autograd Function wiring against the reference's autograd goldens"""
    from tests.cpu_emulation import su3_host_logic_on_cpu
    import tests.test_gpu_training as tg
    monkeypatch.setattr(tg, 'DEV', 'cpu')
    old = torch.get_default_dtype()
    try:
        with su3_host_logic_on_cpu(monkeypatch):
            tg.test_su3_l2hmc_gradients_match_reference_autograd(golden_dir, torch.set_default_dtype)
    finally:
        torch.set_default_dtype(old)

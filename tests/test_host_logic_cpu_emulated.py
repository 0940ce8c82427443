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

"""GPU tier: the un-merged transition kernel (`merge_directions = False`, dynamics.py:1031-1063, with the
reference's swapped states in its accept probability) and the verbose per-step metrics of all three kernels
(`get_metrics` / `update_history`, dynamics.py:865-898) against goldens frozen from the reference.
(File name sorts last on purpose: it was added after the round's last GPU run.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.detach().cpu().numpy()


def maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


def _dynamics(gu, verbose):
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    shape, nb, nlf = [int(s) for s in gu['shape']], 3, 2
    pre = 'dense/'
    cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=nlf, eps=0.1, eps_hmc=0.1, use_ncp=True,
                         verbose=verbose, use_split_xnets=True, use_separate_networks=True, merge_directions=True)
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=[16, 12], activation_fn='leaky_relu', dropout_prob=0.2,
                                                      use_batch_norm=True),
                         conv_config=None, net_weights=None)
    lat = LatticeU1(nb, shape)
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    sd = {k[len(pre) + 3:]: torch.from_numpy(gu[k]) for k in gu.files if k.startswith(pre + 'sd/')}
    dyn.load_state_dict(sd, strict=False)
    dyn.masks = [dev(m) for m in gu[pre + 'masks']]
    dyn.eval()
    return dyn


@pytest.mark.parametrize('tag,tol', [('f64', 1e-11), ('f32', 2e-5)])
def test_unmerged_transition_kernel_matches_reference(golden_dir, tag, tol):
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64 if tag == 'f64' else torch.float32)
    try:
        gu = np.load(golden_dir / f'u1_{tag}.npz')
        pre = 'dense/'
        dyn = _dynamics(gu, verbose=False)
        st = State(dev(gu['x']), dev(gu[pre + 'v']), torch.tensor(float(gu['beta'])))
        for key, fwd in (('tkf', True), ('tkb', False)):
            with torch.no_grad():
                sp, met = dyn.transition_kernel(st, forward=fwd)
            assert maxdiff(host(sp.x).reshape(gu[f'{pre}{key}_x'].shape), gu[f'{pre}{key}_x']) <= 50 * tol
            assert maxdiff(host(sp.v).reshape(3, -1), gu[f'{pre}{key}_v']) <= 50 * tol
            assert maxdiff(host(met['sumlogdet']), gu[f'{pre}{key}_sumlogdet']) <= 50 * tol
            assert maxdiff(host(met['acc']), gu[f'{pre}{key}_acc']) <= 200 * tol
    finally:
        torch.set_default_dtype(old)


@pytest.mark.parametrize('tag,tol', [('f64', 1e-11), ('f32', 2e-5)])
def test_verbose_histories_match_reference(golden_dir, tag, tol):
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64 if tag == 'f64' else torch.float32)
    try:
        gu = np.load(golden_dir / f'u1_{tag}.npz')
        pre = 'dense/'
        dyn = _dynamics(gu, verbose=True)
        st = State(dev(gu['x']), dev(gu[pre + 'v']), torch.tensor(float(gu['beta'])))
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
    finally:
        torch.set_default_dtype(old)


def test_forward_without_merge_directions_goes_through_apply_transition(golden_dir):
    """`Dynamics.forward` with merge_directions = False (dynamics.py:616-627,704-742): one random direction,
    accept/reject mix, the metrics contract; every output chain is either its proposal or its input"""
    from l2hmc_b200.dynamics.pytorch.dynamics import State  # noqa: F401
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        gu = np.load(golden_dir / 'u1_f64.npz')
        dyn = _dynamics(gu, verbose=False)
        dyn.config.merge_directions = False
        x, beta = dev(gu['x']), torch.tensor(float(gu['beta']))
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
        assert np.array_equal(host(met['sumlogdet']) != 0, (ma == 1.0) & (host(met['sumlogdet']) != 0))
    finally:
        torch.set_default_dtype(old)

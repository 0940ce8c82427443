"""GPU tier (-m gpu): the `Dynamics` / `Lattice` / `NetworkFactory` mirrors (the
drop-in boundary) driven exactly like the reference's trainer drives
its own classes, against golden outputs of the reference's `Dynamics`:
the reference's weights (`state_dict`), masks and step sizes are loaded into our
module tree, so key names and shapes are checked on the way.

L2HMC SU(3) tolerance is 1e-8, not 1e-12: the reference feeds the anti-Hermitian
FORCE through projectSU (dynamics.py:1154-1156), which is ill conditioned -- two
correct evaluations (reference vs numpy oracle) already differ by 3e-10 there
(tests/test_oracle_golden.py::test_su3_group_ops)."""
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
    return float(np.max(np.abs(np.asarray(a, dtype=np.complex128) - np.asarray(b, dtype=np.complex128))))


@pytest.fixture()
def default_dtype():
    old = torch.get_default_dtype()
    yield torch.set_default_dtype
    torch.set_default_dtype(old)


def _load_reference_weights(dyn, gl, prefix=''):
    sd = {k[len(prefix) + 3:]: torch.from_numpy(gl[k]) for k in gl.files if k.startswith(prefix + 'sd/')}
    res = dyn.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys, f'reference keys we do not have: {res.unexpected_keys[:5]}'
    # every missing key must be the reference's duplicated `networks.` alias of a loaded one
    # (or the dead SU(3) xnet, which the golden file does not carry)
    for k in res.missing_keys:
        assert k.startswith('networks.') or k.startswith('xnet.'), k
    return sd


def test_su3_l2hmc_matches_reference(golden_dir, default_dtype):
    default_dtype(torch.float64)
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, NetWeights, NetWeight, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    gl = np.load(golden_dir / 'su3_l2hmc_f64.npz')
    shape, nb, nlf = [int(s) for s in gl['shape']], 2, int(gl['nlf'])
    cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=nlf, eps=0.05, eps_hmc=0.1,
                         verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=[8], activation_fn='tanh', dropout_prob=0.0,
                                                      use_batch_norm=False),
                         conv_config=None, net_weights=NetWeights(x=NetWeight(0., 1., 1.), v=NetWeight(1., 1., 1.)))
    lat = LatticeSU3(nb, shape)
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    _load_reference_weights(dyn, gl)
    dyn.masks = [dev(m) for m in gl['masks']]
    dyn.eval()
    assert dyn.vnet.input_layer.xlayer.weight.shape == (8, 4 * int(np.prod(shape)) * 8)
    st = State(dev(gl['x']), dev(gl['v']), torch.tensor(float(gl['beta'])))
    with torch.no_grad():
        s1, ld = dyn._update_v_fwd(0, st)
        assert maxdiff(host(s1.v), gl['vfwd_v']) < 1e-8 and maxdiff(host(ld), gl['vfwd_logdet']) < 1e-8
        s1, ld = dyn._update_v_bwd(1, st)
        assert maxdiff(host(s1.v), gl['vbwd_v']) < 1e-8 and maxdiff(host(ld), gl['vbwd_logdet']) < 1e-8
        m, mb = dyn._get_mask(0)
        s2, ld = dyn._update_x_fwd(0, st, m, first=True)
        assert maxdiff(host(s2.x), gl['xfwd_x']) < 1e-12 and float(ld.abs().max()) == 0
        s2, _ = dyn._update_x_bwd(0, st, m, first=True)
        assert maxdiff(host(s2.x), gl['xbwd_x']) < 1e-12
        sp, met = dyn.transition_kernel_fb(st)
    assert maxdiff(host(sp.x), gl['fb_x']) < 1e-8
    assert maxdiff(host(sp.v), gl['fb_v']) < 1e-8
    assert maxdiff(host(met['acc']), gl['fb_acc']) < 1e-8
    assert maxdiff(host(met['sumlogdet']), gl['fb_sumlogdet']) < 1e-8
    # the full public call: shapes / dtypes of the metrics contract
    with torch.no_grad():
        xout, metrics = dyn((st.x, st.beta))
    assert xout.shape == (nb, dyn.xdim) and xout.dtype == torch.complex128
    assert metrics['acc'].shape == (nb,) and metrics['acc_mask'].dtype == torch.float32
    assert metrics['sumlogdet'].shape == (nb,)
    mc = metrics['mc_states']
    assert mc.init.x.shape == st.x.shape and mc.out.x.shape == (nb, dyn.xdim)
    # with autograd enabled the same call builds a graph (training path)
    xout2, metrics2 = dyn((st.x, st.beta))
    assert metrics2['acc'].requires_grad


@pytest.mark.parametrize('tag,tol', [('f64', 1e-11), ('f32', 2e-5)])
@pytest.mark.parametrize('name', ['dense', 'conv'])
@pytest.mark.parametrize('fused', ['auto', 'never'])      # heads fused with the update (l2b_u1_heads_update) or not
def test_u1_l2hmc_matches_reference(golden_dir, default_dtype, tag, tol, name, fused):
    default_dtype(torch.float64 if tag == 'f64' else torch.float32)
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, ConvolutionConfig, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    shape, nb, nlf = [int(s) for s in gu['shape']], 3, 2
    pre = f'{name}/'
    cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=nlf, eps=0.1, eps_hmc=0.1, use_ncp=True,
                         verbose=False, use_split_xnets=True, use_separate_networks=True, merge_directions=True)
    conv = ConvolutionConfig(filters=[4, 8, 8], sizes=[3, 2, 2], pool=[2, 2, 2]) if name == 'conv' else None
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=[16, 12], activation_fn='leaky_relu', dropout_prob=0.2,
                                                      use_batch_norm=True),
                         conv_config=conv, net_weights=None)
    lat = LatticeU1(nb, shape)
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    sd = _load_reference_weights(dyn, gu, pre)
    ours = {k for k in dyn.state_dict() if not k.startswith('networks.')}
    assert ours == set(sd), f'state_dict key mismatch: {sorted(ours ^ set(sd))[:6]}'
    dyn.masks = [dev(m) for m in gu[pre + 'masks']]
    dyn.eval()
    dyn.fused_u1_heads = fused
    from l2hmc_b200 import _lib
    n0 = _lib.launch_count()
    st = State(dev(gu['x']), dev(gu[pre + 'v']), torch.tensor(float(gu['beta'])))
    with torch.no_grad():
        m, _ = dyn._get_mask(0)
        s2, ld2 = dyn._update_x_fwd(0, st, m, first=True)
        assert maxdiff(host(s2.x), gu[pre + 'xfwd_x']) <= 10 * tol
        assert maxdiff(host(ld2), gu[pre + 'xfwd_logdet']) <= 10 * tol
        s3, ld3 = dyn._update_x_bwd(1, st, m, first=False)
        assert maxdiff(host(s3.x), gu[pre + 'xbwd_x']) <= 10 * tol
        assert maxdiff(host(ld3), gu[pre + 'xbwd_logdet']) <= 10 * tol
        s4, ld4 = dyn._update_v_fwd(0, st)
        assert maxdiff(host(s4.v), gu[pre + 'vfwd_v']) <= 10 * tol
        assert maxdiff(host(ld4), gu[pre + 'vfwd_logdet']) <= 10 * tol
        sp, met = dyn.transition_kernel_fb(st)
    assert maxdiff(host(sp.x), gu[pre + 'fb_x']) <= 50 * tol
    assert maxdiff(host(sp.v).reshape(nb, -1), gu[pre + 'fb_v']) <= 50 * tol
    assert maxdiff(host(met['acc']), gu[pre + 'fb_acc']) <= 50 * tol
    assert maxdiff(host(met['sumlogdet']), gu[pre + 'fb_sumlogdet']) <= 50 * tol


def test_hmc_public_api_contract(golden_dir, default_dtype):
    """apply_transition_hmc: proposal == reference golden for injected momenta,
    output = per-chain select, metrics keys as the trainer reads them"""
    default_dtype(torch.float64)
    from l2hmc_b200.configs import DynamicsConfig
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    g = np.load(golden_dir / 'su3_f64.npz')
    shape, nb = [int(s) for s in g['shape']], 2
    for verbose in (False, True):
        cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=2, eps=0.05, eps_hmc=0.05,
                             verbose=verbose, use_split_xnets=False, use_separate_networks=False)
        lat = LatticeSU3(nb, shape)
        dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
        st = State(dev(g['xw']), dev(g['vw']), torch.tensor(float(g['beta'])))
        sp, met = dyn.transition_kernel_hmc(st, eps=float(g['hmcw_eps']), nleapfrog=int(g['hmcw_nlf']))
        assert maxdiff(host(sp.x), g['hmcw_x']) < 1e-12 and maxdiff(host(sp.v), g['hmcw_v']) < 1e-12
        assert maxdiff(host(met['acc']), g['hmcw_acc']) < 1e-12 * max(1.0, float(np.abs(g['hmcw_h0']).max()))
        if verbose:
            assert met['energy'].shape == (int(g['hmcw_nlf']) + 1, nb)
            assert np.allclose(host(met['energy'][0]), g['hmcw_h0'], rtol=1e-12)
            assert np.allclose(host(met['energy'][-1]), g['hmcw_h1'], rtol=1e-12)
    torch.manual_seed(3)
    xout, metrics = dyn.apply_transition_hmc((st.x, st.beta))      # merge_directions -> 2*nlf steps
    mc = metrics['mc_states']
    ma = metrics['acc_mask']
    assert ma.dtype == torch.float32 and set(ma.tolist()) <= {0.0, 1.0}
    want = torch.where(ma[:, None].bool(), mc.proposed.x.flatten(1), mc.init.x.flatten(1))
    assert torch.equal(xout, want) and xout.shape == (nb, dyn.xdim)
    p = mc.init.v
    assert float((p + p.adjoint()).abs().max()) < 1e-15, 'fresh momenta must be anti-Hermitian'
    # lattice observables through the mirror == golden
    m = lat.calc_metrics(dev(g['x']), beta=torch.tensor(float(g['beta'])))
    assert maxdiff(host(m['plaqs']), g['plaqs']) < 1e-13 and maxdiff(host(m['intQ']), g['intQ']) < 1e-13
    assert np.allclose(host(m['action']), g['action'], rtol=1e-12, atol=1e-12)
    assert maxdiff(host(m['dsdx']), g['force']) < 1e-12
    assert maxdiff(host(lat.g.group_to_vec(dev(g['x']))), g['vec_x']) < 1e-12
    assert maxdiff(host(lat.g.update_gauge(dev(g['x']), 0.1 * dev(g['v']))), g['upd']) < 1e-12


def test_hmc_half_updates_and_apply_transition_both(default_dtype):
    """the remaining public Dynamics methods of the reference (dynamics.py:744-803,1244-1264)"""
    from l2hmc_b200 import ops
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    default_dtype(torch.float64)
    torch.manual_seed(0)
    np.random.seed(0)
    # U(1)
    nb, shape = 5, [6, 4]
    cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=2, eps=0.2, verbose=False,
                         merge_directions=False)
    fac = NetworkFactory(input_spec=get_input_spec(cfg), network_config=NetworkConfig(units=[8], activation_fn='tanh',
                         dropout_prob=0.0, use_batch_norm=False), conv_config=None, net_weights=None)
    lat = LatticeU1(nb, shape)
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    x, v, beta = lat.random(), lat.g.random_momentum([nb, 2, *shape]), torch.tensor(2.5)
    st = State(x, v, beta)
    e = float(torch.sigmoid(dyn.veps[1].log()))
    f = lat.grad_action(x, beta).reshape(nb, -1)
    assert float((dyn._update_v_fwd_hmc(1, st) - (v - 0.5 * e * f)).abs().max()) < 1e-13
    assert float((dyn._update_v_bwd_hmc(1, st) - (v + 0.5 * e * f)).abs().max()) < 1e-13
    ex = float(torch.sigmoid(dyn.xeps[0].log()))
    assert float((dyn._update_x_fwd_hmc(0, st) - (x.reshape(nb, -1) + ex * v)).abs().max()) < 1e-13
    assert float((dyn._update_x_bwd_hmc(0, st) - (x.reshape(nb, -1) - ex * v)).abs().max()) < 1e-13
    with torch.no_grad():
        xo, m = dyn.apply_transition_both((x, beta))
    ms = m['mc_states']
    assert xo.shape == (nb, 2 * 24) and m['acc'].shape == (nb,) and set(m['acc_mask'].tolist()) <= {0.0, 1.0}
    for b in range(nb):      # per chain: out is the proposal iff accepted
        want = ms.proposed.x[b] if float(m['acc_mask'][b]) == 1.0 else x[b].reshape(-1)
        assert torch.equal(xo[b], want.reshape(-1))
    z = torch.randn(3, 2, 4, device=DEV)
    assert torch.equal(Dynamics.complexify(z), torch.complex(z[:, 0], z[:, 1]))
    assert dyn._stack_as_xy(x).shape == (*x.shape, 2)
    # SU(3)
    nb, shape = 2, [2, 2, 2, 4]
    cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=1, eps=0.1, verbose=False,
                         use_split_xnets=False, use_separate_networks=False)
    lat3 = LatticeSU3(nb, shape)
    dyn3 = Dynamics(potential_fn=lat3.action, config=cfg, network_factory=None)
    x3, v3 = lat3.random(), lat3.random_momentum()
    st3 = State(x3, v3, torch.tensor(5.5))
    e3 = float(torch.sigmoid(dyn3.xeps[0].log()))
    assert float((dyn3._update_x_fwd_hmc(0, st3) - ops.su3_update_gauge(x3, v3, e3)).abs().max()) < 1e-14
    assert float((dyn3._update_x_bwd_hmc(0, st3) - ops.su3_update_gauge(x3, v3, -e3)).abs().max()) < 1e-14
    ev = float(torch.sigmoid(dyn3.veps[0].log()))
    f3 = lat3.grad_action(x3, torch.tensor(5.5))
    assert float((dyn3._update_v_fwd_hmc(0, st3) - (v3 - 0.5 * ev * f3)).abs().max()) < 1e-13


@pytest.mark.parametrize('tag,shape,tol', [('f64', [96, 80], 1e-11), ('f32', [128, 112], 2e-4)])
def test_u1_hmc_beyond_one_block_of_shared_memory_and_zero_steps(default_dtype, tag, shape, tol):
    """The whole-trajectory U(1) kernel keeps 5 T X elements of one chain in shared memory; lattices beyond 227 KB
    (96x80 in f64, 128x112 in f32) take the per-step kernels instead of raising, and `nleapfrog=0` returns the
    state unchanged with acc = 1 -- both as the reference behaves (dynamics.py:900-954)."""
    from l2hmc_b200.configs import DynamicsConfig
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from oracle import dynamics as od
    tdt, ndt = (torch.float64, np.float64) if tag == 'f64' else (torch.float32, np.float32)
    default_dtype(tdt)
    nb, nlf, eps, beta = 3, 4, 0.05, 2.5
    rng = np.random.default_rng(17)
    x = rng.uniform(-np.pi, np.pi, (nb, 2, *shape)).astype(ndt)
    v = rng.standard_normal((nb, 2 * shape[0] * shape[1])).astype(ndt)
    cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=nlf, eps=eps, eps_hmc=eps, verbose=False)
    lat = LatticeU1(nb, shape)
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
    st = State(dev(x), dev(v), torch.tensor(beta))
    prop, met = dyn.transition_kernel_hmc(st, eps=eps, nleapfrog=nlf)
    want, acc = od.transition_kernel_hmc(od.U1Ops, od.State(x.astype(np.float64), v.astype(np.float64), beta), eps, nlf)
    assert maxdiff(host(prop.x).reshape(nb, -1), want.x.reshape(nb, -1)) < tol
    assert maxdiff(host(prop.v).reshape(nb, -1), want.v.reshape(nb, -1)) < tol
    assert maxdiff(host(met['acc']), acc) < (1e-9 if tag == 'f64' else 5e-2)
    same, met0 = dyn.transition_kernel_hmc(st, eps=eps, nleapfrog=0)
    assert torch.equal(same.x.reshape(nb, -1), st.x.reshape(nb, -1)) and torch.equal(same.v, st.v)
    assert torch.equal(met0['acc'], torch.ones_like(met0['acc']))


def test_checkpoint_save_load_round_trip(tmp_path):
    """Dynamics.save / load / save_eps / restore_eps (dynamics.py:544-586): a second Dynamics built from other seeds
    reproduces the first one's sweep bit for bit after loading its checkpoint; the step sizes survive the .npy files"""
    from tests._helpers import _su3_trainer
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        torch.manual_seed(1)
        np.random.seed(1)
        tr1, lat = _su3_trainer(nb=2, units=(8,))
        torch.manual_seed(2)
        np.random.seed(1)                                 # same leapfrog masks (numpy), different weights (torch)
        tr2, _ = _su3_trainer(nb=2, units=(8,))
        d1, d2 = tr1.dynamics, tr2.dynamics
        with torch.no_grad():
            for p in list(d1.xeps) + list(d1.veps):
                p.mul_(1.37)
        x = lat.random().to(torch.complex128)
        v = lat.random_momentum()
        st = State(x, v, torch.tensor(6.0))
        d1.eval()
        d2.eval()
        with torch.no_grad():
            a, ma = d1.transition_kernel_fb(st)
            b0, _ = d2.transition_kernel_fb(st)
        assert float((a.x - b0.x).abs().max()) > 1e-6            # different weights: different proposal
        d1.save(tmp_path)
        d2.load(tmp_path)
        with torch.no_grad():
            b, mb = d2.transition_kernel_fb(st)
        assert torch.equal(a.x, b.x) and torch.equal(a.v, b.v) and torch.equal(ma['acc'], mb['acc'])
        with torch.no_grad():
            for p in list(d2.xeps) + list(d2.veps):
                p.fill_(0.5)
        d2.restore_eps(tmp_path)
        for p1, p2 in zip(list(d1.xeps) + list(d1.veps), list(d2.xeps) + list(d2.veps)):
            assert float((p1 - p2).abs().max()) == 0.0
    finally:
        torch.set_default_dtype(old)

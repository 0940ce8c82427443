"""GPU tier (-m gpu): the CUDA U(1) path through the C ABI against the golden
vectors of the reference (fp32 and fp64) and the numpy oracle.

Tolerances: fp64 1e-12; fp32 1e-5 (north star), applied relative to the
magnitude of per-chain sums.  "Bit-exact integer topological charge" is checked
as round(intQ) equality (intQ itself is a float sum of wrapped angles / 2 pi,
SURVEY 7.3)."""
import numpy as np
import pytest
import torch

from oracle import u1 as ou1, dynamics as od

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.detach().cpu().numpy()


def maxdiff(a, b):
    return float(np.max(np.abs(np.asarray(a, dtype=np.float64) - np.asarray(b, dtype=np.float64))))


@pytest.fixture(scope='module')
def ops():
    from l2hmc_b200 import ops as _ops
    return _ops


@pytest.mark.parametrize('tag,tol', [('f64', 1e-12), ('f32', 1e-5)])
def test_u1_lattice_ops(ops, golden_dir, tag, tol):
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    x, beta = dev(gu['x']), float(gu['beta'])
    assert maxdiff(host(ops.u1_wilson_loops(x)), gu['wloops']) <= tol
    obs = host(ops.u1_observables(x, beta))
    assert obs.dtype == gu['x'].dtype
    assert np.allclose(obs[:, 0], gu['action'], rtol=tol)
    assert maxdiff(obs[:, 1], gu['plaqs']) <= tol
    assert maxdiff(obs[:, 2], gu['sinQ']) <= tol
    assert maxdiff(obs[:, 3], gu['intQ']) <= tol
    assert np.array_equal(np.round(obs[:, 3]), np.round(gu['intQ'])), 'integer topological charge must match exactly'
    assert maxdiff(host(ops.u1_force(x, beta)), gu['force']) <= 10 * tol
    assert maxdiff(host(ops.u1_compat_proj(3.0 * x)), gu['compat']) <= 10 * tol
    assert np.allclose(host(ops.u1_kinetic(dev(gu['v']))), ou1.kinetic_energy(gu['v']), rtol=tol)


@pytest.mark.parametrize('tag,tol', [('f64', 1e-12), ('f32', 1e-5)])
def test_u1_hmc_trajectory_vs_reference_golden(ops, golden_dir, tag, tol):
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    shape = [int(s) for s in gu['shape']]
    x, v, beta = dev(gu['x']), dev(gu['v']), float(gu['beta'])
    xo, vo, en = ops.u1_hmc_trajectory(x, v, beta, 0.1, 5, shape=shape)
    assert maxdiff(host(xo).reshape(3, -1), gu['hmc_x']) <= 10 * tol
    assert maxdiff(host(vo).reshape(3, -1), gu['hmc_v']) <= 10 * tol
    en = host(en).astype(np.float64)
    h0, h1 = en[:, 0] + en[:, 1], en[:, 2] + en[:, 3]
    assert np.allclose(h0, gu['hmc_h0'], rtol=tol) and np.allclose(h1, gu['hmc_h1'], rtol=tol)
    acc = np.exp(np.minimum(h0 - h1, 0))
    assert maxdiff(acc, gu['hmc_acc']) <= tol * max(1.0, float(np.abs(gu['hmc_h0']).max()))


@pytest.mark.parametrize('tag,tol', [('f64', 1e-12), ('f32', 1e-5)])
def test_u1_l2hmc_elementwise_updates(ops, golden_dir, tag, tol):
    """x/v update kernels fed with the (s, t, q) the oracle's networks produce"""
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    shape = [int(s) for s in gu['shape']]
    pre = 'dense/'
    sd = {k[len(pre) + 3:]: gu[k] for k in gu.files if k.startswith(pre + 'sd/')}
    spec = od.L2HMCSpec(group='U1', xshape=(3, 2, *shape), nleapfrog=2, xeps=list(gu[pre + 'xeps']),
                        veps=list(gu[pre + 'veps']), masks=list(gu[pre + 'masks']), state_dict=sd,
                        activation='leaky_relu', use_batch_norm=True)
    x, v, beta = gu['x'], gu[pre + 'v'], float(gu['beta'])
    dt = x.dtype.type
    m = spec.masks[0]
    for step, first, sign, kx, kl, mask in ((0, True, +1, 'xfwd_x', 'xfwd_logdet', m),
                                            (1, False, -1, 'xbwd_x', 'xbwd_logdet', m)):
        eps = od._sig_log(spec.xeps[step], x.dtype)
        s, t, q = od.call_xnet(spec, step, m.reshape(1, 2, *shape) * x, v, first)
        xo, ld = ops.u1_xupdate(dev(x), dev(v), dev(s), dev(t), dev(q), dev(mask), float(eps), sign, True)
        assert maxdiff(host(xo), gu[pre + kx]) <= 20 * tol
        assert maxdiff(host(ld), gu[pre + kl]) <= 20 * tol
    f = ou1.grad_action(x, beta)
    s, t, q = od.call_vnet(spec, 0, x, f)
    eps = od._sig_log(spec.veps[0], x.dtype)
    vo, ld = ops.u1_vupdate(dev(v), dev(f), dev(s), dev(t), dev(q), float(eps), +1)
    assert maxdiff(host(vo), gu[pre + 'vfwd_v']) <= 20 * tol
    assert maxdiff(host(ld), gu[pre + 'vfwd_logdet']) <= 20 * tol
    # non-NCP affine update vs the oracle
    spec.use_ncp = False
    st = od.State(x, v, beta)
    for sign, fwd in ((+1, True), (-1, False)):
        s2, ld2 = od.update_x(spec, 0, st, m, True, fwd)
        sn, tn, qn = od.call_xnet(spec, 0, m.reshape(1, 2, *shape) * x, v, True)
        eps = od._sig_log(spec.xeps[0], x.dtype)
        xo, ld = ops.u1_xupdate(dev(x), dev(v), dev(sn), dev(tn), dev(qn), dev(m), float(eps), sign, False)
        assert maxdiff(host(xo), s2.x) <= 20 * tol and maxdiff(host(ld), ld2) <= 20 * tol


@pytest.mark.parametrize('T,X,nb,dtype', [(2, 2, 1, np.float64), (16, 16, 128, np.float32), (7, 5, 3, np.float64),
                                          (64, 64, 16, np.float32), (64, 64, 4, np.float64)])
def test_u1_vs_oracle_shapes(ops, T, X, nb, dtype):
    rng = np.random.default_rng(T * X + nb)
    x = rng.uniform(-np.pi, np.pi, (nb, 2, T, X)).astype(dtype)
    v = rng.standard_normal((nb, 2 * T * X)).astype(dtype)
    tol = 1e-12 if dtype == np.float64 else 1e-5
    beta, eps, nlf = 4.0, 0.05, 10
    s, acc = od.transition_kernel_hmc(od.U1Ops, od.State(x, v, beta), eps, nlf)
    xo, vo, en = ops.u1_hmc_trajectory(dev(x), dev(v), beta, eps, nlf, shape=[T, X])
    assert maxdiff(host(xo).reshape(nb, -1), s.x) <= 20 * tol
    assert maxdiff(host(vo).reshape(nb, -1), s.v) <= 20 * tol
    en = host(en).astype(np.float64)
    h0 = (ou1.kinetic_energy(v) + ou1.action(x, beta)).astype(np.float64)
    assert np.allclose(en[:, 0] + en[:, 1], h0, rtol=10 * tol)
    obs = host(ops.u1_observables(dev(x), beta))
    assert np.array_equal(np.round(obs[:, 3]), np.round(ou1.int_charges(x.astype(np.float64))))


def test_u1_reversibility_full_size(ops):
    """BASELINE cfg 2 shape (64x64, 4096 chains, fp32): run forward, flip v, run back"""
    nb, T, X = 4096, 64, 64
    torch.manual_seed(1)
    x = (torch.rand(nb, 2, T, X, device=DEV) * 2 - 1) * np.pi
    v = torch.randn(nb, 2, T, X, device=DEV)
    x1, v1, en = ops.u1_hmc_trajectory(x, v, 4.0, 0.1, 10)
    x2, v2, _ = ops.u1_hmc_trajectory(x1, -v1, 4.0, 0.1, 10)
    assert float((x2 - x).abs().max()) < 5e-4     # fp32 round-off over 20 steps of O(10) angles
    assert float((v2 + v).abs().max()) < 5e-4
    assert torch.isfinite(en).all()


def test_fused_u1_sweep_is_as_accurate_as_the_unfused_one():
    """fp32 L2HMC sweep with the fused input-layer / heads kernels vs the unfused fp32 path, both measured
    against a float64 evaluation of the same nets: the log-Jacobian is a sum over the lattice, so a biased
    1e-7 per element (SFU intrinsics in the wrong place) shows up as 1e-2 per chain -- it must not."""
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    old = torch.get_default_dtype()
    nb, shape, nlf = 16, [32, 32], 4

    def build(dtype):
        torch.set_default_dtype(dtype)
        torch.manual_seed(1)
        np.random.seed(1)
        cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=nlf, eps=0.1, use_ncp=True,
                             verbose=False, use_split_xnets=True, merge_directions=True, use_separate_networks=True)
        fac = NetworkFactory(input_spec=get_input_spec(cfg),
                             network_config=NetworkConfig(units=[16, 16], activation_fn='leaky_relu', dropout_prob=0.2,
                                                          use_batch_norm=True), conv_config=None, net_weights=None)
        lat = LatticeU1(nb, shape)
        dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
        dyn.eval()
        return dyn, lat
    try:
        d32, lat32 = build(torch.float32)
        d64, _ = build(torch.float64)
        d64.load_state_dict(d32.state_dict())
        d64.masks = [m.clone() for m in d32.masks]
        torch.manual_seed(5)
        x = lat32.random().float()
        v = torch.randn(nb, 2 * 32 * 32, device=x.device, dtype=torch.float32)
        beta = torch.tensor(4.0)
        err = {}
        with torch.no_grad():
            ref, mref = d64.transition_kernel_fb(State(x.double(), v.double(), beta))
            for mode in ('never', 'auto'):
                d32.fused_u1_heads = mode
                st, met = d32.transition_kernel_fb(State(x, v, beta))
                err[mode] = (float((met['sumlogdet'].double() - mref['sumlogdet']).abs().max()),
                             float((met['acc'].double() - mref['acc']).abs().max()),
                             float((st.v.double().reshape(nb, -1) - ref.v.reshape(nb, -1)).abs().max()))
        for k in range(3):
            assert err['auto'][k] <= 3.0 * err['never'][k] + 1e-6, (k, err)
    finally:
        torch.set_default_dtype(old)


@pytest.mark.parametrize('dtype', [torch.float32, torch.float64])
def test_u1_wilson_loops4x4_kernel_equals_the_reference(dtype):
    """LatticeU1.wilson_loops4x4 / plaqs4x4 (lattice/u1/pytorch/lattice.py:161-219) on `l2b_u1_wilson_loops4x4`
    against the reference's own rolled sum (oracle/_ref when it travelled) and our torch restatement of it under
    autograd; same summation order, so fp32 agrees to the last bit with the torch path"""
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        for nb, T, X in ((5, 8, 8), (3, 6, 10), (2, 16, 16), (2, 5, 7)):
            lat = LatticeU1(nb, [T, X])
            g = torch.Generator(device='cpu').manual_seed(7)
            x = ((torch.rand(nb, 2, T, X, generator=g, dtype=dtype) - 0.5) * 6.2).cuda()
            got = lat.wilson_loops4x4(x)
            xg = x.clone().requires_grad_(True)
            via_torch = lat.wilson_loops4x4(xg)                      # autograd branch: the rolled torch ops
            assert got.shape == via_torch.shape == (X, T, nb)
            assert torch.equal(got, via_torch.detach())
            assert float((lat.plaqs4x4(x) - via_torch.detach().cos().mean((1, 2))).abs().max()) == 0.0
            from oracle import ref_shim
            if ref_shim.available():
                try:
                    ref = ref_shim.load_reference(dtype)
                except RuntimeError:                     # one default dtype per process for the reference copy
                    continue
                rl = ref.LatticeU1(nb, [T, X])
                want = rl.wilson_loops4x4(x.cpu())
                assert float((got.cpu() - want).abs().max()) <= (1e-5 if dtype == torch.float32 else 1e-13)
    finally:
        torch.set_default_dtype(old)

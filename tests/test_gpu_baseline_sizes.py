"""GPU tier (-m gpu): parity AT the BASELINE lattice sizes, not only on the small golden shapes.

The CUDA trajectory (C ABI `l2b_su3_hmc_trajectory`, and the public `Dynamics.transition_kernel_hmc`) is
compared on identical seeded `(x, v, beta)` with

* the reference's own `Dynamics.transition_kernel_hmc` (`dynamics/pytorch/dynamics.py:915-954`, `leapfrog_hmc`
  `:900-913`; action `lattice/su3/pytorch/lattice.py:252-269`, force by autograd `:299-308`) run in the test from
  the unmodified copy `oracle/_ref` (it travels with the snapshot) -- on whatever device the reference picks for
  itself: it moves its module-level constants to CUDA as soon as torch sees a GPU (`l2hmc/__init__.py:45-51`,
  `dynamics.py:194-200`), so on the GPU box this is the reference's own ATen/CUDA path, not ours -- and
* the numpy oracle (`oracle/dynamics.py`), which needs nothing but numpy and so also runs where `oracle/_ref`
  did not travel.

Sizes: 8^4 x 2 chains x N_LF 10 (the lattice of BASELINE cfg 3 / 5) and 16^4 x 1 chain x N_LF 3 (the lattice of
cfg 4, the headline).  Tolerances as tests/test_gpu_su3.py: links and momenta 1e-12 absolute, action / H 1e-12
relative, acc within 1e-12 * max(1, |H|).
"""
import numpy as np
import pytest
import torch

from oracle import dynamics as od
from oracle import ref_shim
from oracle import su3 as osu3

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'

CASES = [([8, 8, 8, 8], 2, 10, 0.1, 6.0), ([16, 16, 16, 16], 1, 3, 0.1, 6.0)]


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def host(t):
    return t.detach().cpu().numpy()


def _inputs(shape, nb):
    rng = np.random.default_rng(1000 + shape[0] + nb)
    full = (nb, 4, *shape, 3, 3)
    x, v = osu3.random_su3(rng, full), osu3.random_momentum(rng, full)
    if nb > 1:
        # chain 0: hot start (dH >> 0, acc = 1); chain 1: a smooth configuration exp(0.272 P) chosen so that
        # dH ~ -1 over this trajectory, i.e. an accept probability strictly inside (0, 1)
        x[1] = osu3.expm(0.272 * osu3.random_momentum(rng, full[1:]))
    return x, v


def _cuda_trajectory(x, v, beta, eps, nlf):
    from l2hmc_b200 import ops
    xo, vo, en = ops.su3_hmc_trajectory(dev(x), dev(v), beta, eps, nlf)
    en = host(en)
    return host(xo), host(vo), en[:, 0] + en[:, 1], en[:, 2] + en[:, 3], en


def _check(got, want, tag):
    gx, gv, gh0, gh1 = got
    wx, wv, wh0, wh1 = want
    assert np.abs(gx - wx).max() < 1e-12, tag
    assert np.abs(gv - wv).max() < 1e-12, tag
    assert np.allclose(gh0, wh0, rtol=1e-12, atol=0), (tag, gh0, wh0)
    assert np.allclose(gh1, wh1, rtol=1e-12, atol=0), (tag, gh1, wh1)
    scale = max(1.0, float(np.abs(wh0).max()))
    ga, wa = np.exp(np.minimum(gh0 - gh1, 0.0)), np.exp(np.minimum(wh0 - wh1, 0.0))
    assert np.abs(ga - wa).max() < 1e-12 * scale, tag
    if len(wa) > 1:
        assert 0.05 < wa[1] < 0.95, wa    # the smooth chain exercises acc inside (0, 1)


@pytest.mark.parametrize('shape,nb,nlf,eps,beta', CASES, ids=['8^4x2_nlf10', '16^4x1_nlf3'])
def test_hmc_trajectory_matches_numpy_oracle_at_baseline_size(shape, nb, nlf, eps, beta):
    x, v = _inputs(shape, nb)
    gx, gv, gh0, gh1, en = _cuda_trajectory(x, v, beta, eps, nlf)
    want, acc = od.transition_kernel_hmc(od.SU3Ops, od.State(x, v, beta), eps, nlf)
    wh0 = osu3.action(x, beta) + osu3.kinetic_energy(v)
    wh1 = osu3.action(want.x, beta) + osu3.kinetic_energy(want.v)
    _check((gx, gv, gh0, gh1), (want.x, want.v, wh0, wh1), 'numpy oracle')
    # energies = (KE0, S0, KE1, S1): the Wilson action and the kinetic energy separately, 1e-12 relative
    assert np.allclose(en[:, 1], osu3.action(x, beta), rtol=1e-12, atol=0)
    assert np.allclose(en[:, 3], osu3.action(want.x, beta), rtol=1e-12, atol=0)
    assert np.allclose(en[:, 0], osu3.kinetic_energy(v), rtol=1e-12, atol=0)
    assert np.allclose(en[:, 2], osu3.kinetic_energy(want.v), rtol=1e-12, atol=0)


@pytest.mark.skipif(not ref_shim.available(), reason='oracle/_ref did not travel')
@pytest.mark.parametrize('shape,nb,nlf,eps,beta', CASES, ids=['8^4x2_nlf10', '16^4x1_nlf3'])
def test_hmc_trajectory_matches_the_reference_itself_at_baseline_size(shape, nb, nlf, eps, beta):
    old = torch.get_default_dtype()
    try:
        ref = ref_shim.load_reference(torch.float64)
        torch.set_default_dtype(torch.float64)
        x, v = _inputs(shape, nb)
        lat = ref.LatticeSU3(nb, shape)
        cfg = ref.DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=nlf, eps=eps, eps_hmc=eps,
                                 verbose=False, use_split_xnets=False, use_separate_networks=False,
                                 merge_directions=True)
        rdyn = ref.Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
        rdev = torch.device('cuda' if torch.cuda.is_available() else 'cpu')   # l2hmc/__init__.py:45-51
        xt, vt, bt = torch.from_numpy(x).to(rdev), torch.from_numpy(v).to(rdev), torch.tensor(beta, device=rdev)
        sp, met = rdyn.transition_kernel_hmc(ref.State(x=xt, v=vt, beta=bt), eps=eps, nleapfrog=nlf)
        wx, wv = host(sp.x).reshape(x.shape), host(sp.v).reshape(v.shape)
        wh0 = host(lat.action(xt, bt) + lat.g.kinetic_energy(vt))
        wh1 = host(lat.action(sp.x.detach().reshape(xt.shape), bt) + lat.g.kinetic_energy(sp.v.detach()))
        racc = host(met['acc'])

        gx, gv, gh0, gh1, _ = _cuda_trajectory(x, v, beta, eps, nlf)
        _check((gx, gv, gh0, gh1), (wx, wv, wh0, wh1), 'reference')
        scale = max(1.0, float(np.abs(wh0).max()))
        assert np.abs(np.exp(np.minimum(gh0 - gh1, 0.0)) - racc).max() < 1e-12 * scale

        # and through the public mirror: same call, same arguments, CUDA underneath
        from l2hmc_b200.configs import DynamicsConfig
        from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
        from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
        mlat = LatticeSU3(nb, shape)
        mcfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=nlf, eps=eps, eps_hmc=eps,
                              verbose=False, use_split_xnets=False, use_separate_networks=False,
                              merge_directions=True)
        mdyn = Dynamics(potential_fn=mlat.action, config=mcfg, network_factory=None)
        mp, mmet = mdyn.transition_kernel_hmc(State(dev(x), dev(v), torch.tensor(beta)), eps=eps, nleapfrog=nlf)
        assert np.abs(host(mp.x).reshape(x.shape) - wx).max() < 1e-12
        assert np.abs(host(mp.v).reshape(v.shape) - wv).max() < 1e-12
        assert np.abs(host(mmet['acc']) - racc).max() < 1e-12 * scale
        # observables of the proposal (`calc_metrics`, lattice.py:310-349): plaquette and both charges
        for name in ('plaqs', 'sinQ', 'intQ'):
            rm = host(lat.calc_metrics(sp.x.detach().reshape(xt.shape))[name])
            mm = host(mlat.calc_metrics(mp.x)[name])
            assert np.allclose(mm, rm, rtol=1e-12, atol=1e-13), name
    finally:
        torch.set_default_dtype(old)

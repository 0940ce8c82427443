"""CPU tier: host-side logic of the `Dynamics` mirror that involves no kernel -- the pieces of
SURVEY appendix B that are pure bookkeeping (masks from numpy's RNG, eps -> eps/(1+eps), the
accept probability and the float32 accept masks).  Methods are called unbound on stand-in objects
because constructing `Dynamics` itself requires a CUDA device (no CPU fallback)."""
import types

import numpy as np
import pytest
import torch

from l2hmc_b200.dynamics.pytorch import dynamics as dmod
from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State


@pytest.mark.parametrize('fname,key,seed', [('u1_f32.npz', 'dense/masks', 3), ('u1_f64.npz', 'conv/masks', 3),
                                            ('su3_l2hmc_f64.npz', 'masks', 7)])
def test_masks_are_the_references_under_the_same_numpy_seed(golden_dir, fname, key, seed):
    """appendix B trap 3: nlf element-wise half masks from np.random.permutation, float32, shape [1, xdim]
    (dynamics.py:1101-1110); the golden generator seeded numpy right before building the reference Dynamics"""
    want = np.load(golden_dir / fname)[key]
    np.random.seed(seed)
    fake = types.SimpleNamespace(config=types.SimpleNamespace(nleapfrog=want.shape[0]), xdim=want.shape[-1])
    got = Dynamics._build_masks(fake)
    assert len(got) == want.shape[0]
    for m, w in zip(got, want):
        assert m.dtype == torch.float32 and tuple(m.shape) == (1, want.shape[-1])
        assert np.array_equal(m.numpy(), w.reshape(1, -1))
        assert int(m.sum()) == want.shape[-1] // 2


def test_l2hmc_step_size_is_eps_over_one_plus_eps():
    """appendix B trap 2 (dynamics.py:82-83,1270,1394)"""
    for e in (0.01, 0.1, 0.5, 2.0):
        p = torch.tensor(e, dtype=torch.float64)
        assert float(dmod.sigmoid(p.log())) == pytest.approx(e / (1 + e), rel=1e-15)
        assert float(Dynamics._eps_t(None, p)) == pytest.approx(e / (1 + e), rel=1e-15)


def test_accept_probability_and_masks():
    """acc = exp(min(H0 - H1 + sumlogdet, 0)); ma = (acc > U[0,1)).float(), mr = 1 - ma, both float32
    (dynamics.py:1065-1087, appendix B trap 10)"""
    h = {'init': torch.tensor([1.0, 2.0, 3.0], dtype=torch.float64), 'prop': torch.tensor([0.5, 2.5, 3.0], dtype=torch.float64)}
    fake = types.SimpleNamespace(_fcache='stale', hamiltonian=lambda st: h[st.x])
    sld = torch.tensor([0.0, 0.2, -0.1], dtype=torch.float64)
    acc = Dynamics.compute_accept_prob(fake, types.SimpleNamespace(x='init'), types.SimpleNamespace(x='prop'), sld)
    want = np.exp(np.minimum(np.array([0.5, -0.3, -0.1]), 0.0))
    assert np.allclose(acc.numpy(), want, rtol=1e-15) and fake._fcache is None
    torch.manual_seed(0)
    px = torch.tensor([0.0, 1.0, 0.5, 1.0], dtype=torch.float64)
    ma, mr = Dynamics._get_accept_masks(px)
    assert ma.dtype == torch.float32 and mr.dtype == torch.float32
    assert ma[0] == 0 and ma[1] == 1 and ma[3] == 1 and torch.equal(ma + mr, torch.ones(4))
    fwd, bwd = Dynamics._get_direction_masks(64)
    assert fwd.dtype == torch.float32 and torch.equal(fwd + bwd, torch.ones(64)) and 8 < int(fwd.sum()) < 56


def test_state_containers():
    """State / MonteCarloStates (dynamics.py:45-74)"""
    x, v = torch.randn(2, 3, 4), torch.randn(2, 12)
    st = State(x, v, torch.tensor(1.5))
    assert st.nb == 2 and tuple(st.xshape) == (2, 3, 4)
    fl = st.flatten()
    assert tuple(fl.x.shape) == (2, 12) and tuple(fl.v.shape) == (2, 12) and fl.beta is st.beta
    d = st.to_numpy()
    assert set(d) == {'x', 'v', 'beta'} and d['x'].shape == (2, 3, 4) and float(d['beta']) == 1.5
    mc = dmod.MonteCarloStates(init=st, proposed=fl, out=fl)
    assert mc.init is st and mc.proposed is fl


def test_hmc_accept_probability_comes_from_potential_fn_unless_it_is_the_kernel_action(monkeypatch):
    """transition_kernel_hmc: the trajectory kernel's own (Wilson) energies are used only when `potential_fn` is
    the plain action of one of our lattices; an improved action (c1 != 0) or any user callable goes through
    `hamiltonian(potential_fn)` like the reference (dynamics.py:1065-1079,1489-1491)"""
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    nb = 2
    en = torch.tensor([[1.0, 2.0, 1.5, 2.5], [1.0, 2.0, 0.5, 1.0]], dtype=torch.float64)

    class FakeOps:
        @staticmethod
        def su3_hmc_trajectory(x, v, beta, eps, nlf):
            return x + 1.0, v + 1.0, en
    monkeypatch.setattr(dmod, 'ops', FakeOps)

    def run(potential_fn):
        fake = types.SimpleNamespace(
            config=types.SimpleNamespace(eps_hmc=0.1, nleapfrog=2, merge_directions=True, verbose=False),
            _su3=True, lattice=types.SimpleNamespace(c1=0.0), potential_fn=potential_fn, _fcache=None,
            _zeros=lambda n: torch.zeros(n, dtype=torch.float64), unflatten=lambda t: t)
        fake._potential_is_wilson = lambda: Dynamics._potential_is_wilson(fake)
        fake.hamiltonian = lambda st: st.x.sum(1) * 0.1       # stands for KE + potential_fn
        fake.compute_accept_prob = lambda a, b, sld: Dynamics.compute_accept_prob(fake, a, b, sld)
        st = State(torch.zeros(nb, 3, dtype=torch.float64), torch.zeros(nb, 3, dtype=torch.float64), torch.tensor(6.0))
        prop, met = Dynamics.transition_kernel_hmc(fake, st)
        assert torch.equal(prop.x, st.x + 1.0) and torch.equal(met['sumlogdet'], torch.zeros(nb, dtype=torch.float64))
        return met['acc'].numpy()
    kernel_acc = np.exp(np.minimum(np.array([3.0 - 4.0, 3.0 - 1.5]), 0.0))
    potential_acc = np.exp(np.minimum(np.array([0.0 - 0.3, 0.0 - 0.3]), 0.0))
    assert np.allclose(run(LatticeSU3(nb, [2, 2, 2, 2]).action), kernel_acc)
    assert np.allclose(run(LatticeSU3(nb, [2, 2, 2, 2], c1=-0.331).action), potential_acc)
    assert np.allclose(run(lambda x, b: x.sum()), potential_acc)
    assert np.allclose(run(LatticeU1(nb, [4, 4]).action), potential_acc)      # a U(1) potential on an SU(3) run
    assert np.allclose(run(LatticeSU3(nb, [2, 2, 2, 2]).kinetic_energy), potential_acc)

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


def test_matrix_valued_loop_fields_of_the_su3_lattice(golden_dir):
    """`_plaquette`, `_trace_plaquette`, `_rectangles`, `_plaquette_field` (lattice.py:93-156): torch code off the
    integrator path, runs on the CPU; traces must reproduce the reference's loops / rectangle goldens"""
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    g = np.load(golden_dir / 'su3_c1_f64.npz')
    shape, nb = [int(s) for s in g['shape']], g['x'].shape[0]
    lat = LatticeSU3(nb, shape, c1=float(g['c1']))
    x = torch.from_numpy(g['x'])
    plaqs, rects = lat._plaquette_field(x.reshape(nb, -1), needs_rect=True)
    assert len(plaqs) == 6 and len(rects) == 12 and tuple(plaqs[0].shape) == (nb, *shape, 3, 3)
    tr = lambda a: torch.diagonal(a, dim1=-2, dim2=-1).sum(-1).numpy()  # noqa: E731
    assert np.abs(np.stack([tr(r) for r in rects]) - g['rects']).max() < 1e-13
    from oracle import su3 as osu3
    assert np.abs(np.stack([tr(p) for p in plaqs]) - osu3.wilson_loops(g['x'])).max() < 1e-13
    assert np.abs(lat._trace_plaquette(x, 2, 1).numpy() - osu3.wilson_loops(g['x'])[2]).max() < 1e-13
    _, zeros = lat._plaquette_field(x)
    assert all(float(z.abs().max()) == 0.0 for z in zeros)
    assert torch.equal(lat._link_staple_op(x[:, 0], x[:, 1]), x[:, 0] @ x[:, 1])
    assert lat.plaq_loss(torch.ones(nb)) is None and lat.charge_loss(torch.ones(nb)) is None   # TODO stubs upstream
    m = dmod.Mask(torch.tensor([1.0, 0.0, 1.0]))
    assert torch.equal(m.combine(torch.tensor([1.0, 2.0, 3.0]), torch.tensor([7.0, 8.0, 9.0])), torch.tensor([1.0, 8.0, 3.0]))


def test_u1_lattice_loss_helpers_and_compat_proju():
    """LatticeU1.plaq_loss / charge_loss / _plaqs4x4 with explicit Wilson loops (lattice.py:205-206,278-308) and
    SU3.compat_proju (group.py:149-165): closed forms, CPU"""
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.group.su3.pytorch.group import SU3
    lat = LatticeU1(2, [4, 4])
    rng = np.random.default_rng(1)
    w1, w2 = torch.from_numpy(rng.uniform(-3, 3, (2, 4, 4))), torch.from_numpy(rng.uniform(-3, 3, (2, 4, 4)))
    acc = torch.tensor([0.25, 1.0], dtype=torch.float64)
    want = -np.mean(acc.numpy() * (2 * (1 - np.cos((w2 - w1).numpy()))).sum((1, 2)) + 1e-4)
    assert float(lat.plaq_loss(acc, wl1=w1, wl2=w2)) == pytest.approx(want, rel=1e-14)
    q1, q2 = np.sin(w1.numpy()).sum((1, 2)) / (2 * np.pi), np.sin(w2.numpy()).sum((1, 2)) / (2 * np.pi)
    assert float(lat.charge_loss(acc, wl1=w1, wl2=w2)) == pytest.approx(-np.mean(acc.numpy() * (q2 - q1) ** 2 + 1e-4), rel=1e-13)
    assert np.allclose(lat._plaqs4x4(w1).numpy(), np.cos(w1.numpy()).mean((1, 2)))
    with pytest.raises(ValueError):
        lat._get_wloops(None)
    u = torch.from_numpy(rng.standard_normal((3, 3, 3)) + 1j * rng.standard_normal((3, 3, 3)))
    xm = torch.from_numpy(rng.standard_normal((3, 3, 3)) + 1j * rng.standard_normal((3, 3, 3)))
    b = SU3().compat_proju(u, xm)
    a0 = np.linalg.solve(u.numpy()[0], xm.numpy()[0])
    b0 = (a0 - a0.conj().T) / 2
    b0 = b0 - np.trace(b0) / 3 * np.eye(3)
    assert np.abs(b.numpy() - b0).max() < 1e-13 and abs(np.trace(b.numpy())) < 1e-13


def test_su3_utils_module_surface():
    """`group/su3/pytorch/utils.py` names used through the reference's import path exist, and the plain-torch
    helpers agree with closed forms / numpy"""
    from l2hmc_b200.group.su3.pytorch import utils as u
    for name in ('projectSU', 'projectU', 'projectTAH', 'rsqrtPHM3', 'rsqrtPHM3f', 'eigs3x3', 'su3_to_vec', 'vec_to_su3',
                 'randTAH3', 'norm2', 'eyeOf', 'checkSU', 'checkU'):
        assert callable(getattr(u, name)), name
    rng = np.random.default_rng(9)
    a = rng.standard_normal((5, 3, 3)) + 1j * rng.standard_normal((5, 3, 3))
    h = a @ a.conj().transpose(0, 2, 1) + 0.5 * np.eye(3)          # positive Hermitian
    ht = torch.from_numpy(h)
    tr = torch.from_numpy(np.trace(h, axis1=1, axis2=2).real)
    p2 = torch.from_numpy(np.trace(h @ h, axis1=1, axis2=2).real)
    det = torch.from_numpy(np.linalg.det(h).real)
    e = torch.stack(u.eigs3x3(tr, p2, det), -1).numpy()
    assert np.allclose(np.sort(e, -1), np.linalg.eigvalsh(h), rtol=1e-10)
    r = u.rsqrtPHM3(ht).numpy()                                     # X^{-1/2}: r h r = 1
    assert np.abs(r @ h @ r - np.eye(3)).max() < 1e-9
    assert torch.equal(u.eyeOf(ht).squeeze().real, torch.eye(3, dtype=u.eyeOf(ht).real.dtype))

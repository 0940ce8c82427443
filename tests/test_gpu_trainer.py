"""GPU tier (-m gpu): `LatticeLoss` values against the reference's own loss on its
own proposal (goldens), and the minimal Trainer step functions (hmc / eval / train)."""
import numpy as np
import pytest
import torch

from tests._helpers import _su3_trainer

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


@pytest.fixture()
def default_dtype():
    old = torch.get_default_dtype()
    yield torch.set_default_dtype
    torch.set_default_dtype(old)


def test_lattice_loss_matches_reference(golden_dir, default_dtype):
    from l2hmc_b200.configs import LossConfig
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.loss.pytorch.loss import LatticeLoss
    default_dtype(torch.float64)
    gl = np.load(golden_dir / 'su3_l2hmc_f64.npz')
    lat = LatticeSU3(2, [int(s) for s in gl['shape']])
    loss = LatticeLoss(lat, LossConfig(use_mixed_loss=False, charge_weight=0.0, rmse_weight=0.1, plaq_weight=0.1))
    got = loss(x_init=dev(gl['x']), x_prop=dev(gl['fb_x']), acc=dev(gl['fb_acc']))
    assert abs(float(got) - float(gl['lattice_loss'])) <= 1e-9 * abs(float(gl['lattice_loss']))
    for tag, tol in (('f64', 1e-10), ('f32', 1e-4)):
        default_dtype(torch.float64 if tag == 'f64' else torch.float32)
        gu = np.load(golden_dir / f'u1_{tag}.npz')
        latu = LatticeU1(3, [int(s) for s in gu['shape']])
        for pre in ('dense/', 'conv/'):
            x0, xp, acc = dev(gu['x']), dev(gu[pre + 'fb_x']), dev(gu[pre + 'fb_acc'])
            l1 = LatticeLoss(latu, LossConfig(use_mixed_loss=True, charge_weight=0.01, rmse_weight=0.0, plaq_weight=0.0))
            l2 = LatticeLoss(latu, LossConfig(use_mixed_loss=False, charge_weight=0.05, rmse_weight=0.0, plaq_weight=0.0))
            assert abs(float(l1(x0, xp, acc)) - float(gu[pre + 'lattice_loss'])) <= tol * abs(float(gu[pre + 'lattice_loss']))
            assert abs(float(l2(x0, xp, acc)) - float(gu[pre + 'lattice_loss2'])) <= tol * abs(float(gu[pre + 'lattice_loss2']))


def test_su3_trainer_steps(default_dtype):
    default_dtype(torch.float64)
    torch.manual_seed(0)
    np.random.seed(0)
    tr, lat = _su3_trainer()
    x = lat.random()
    beta = torch.tensor(6.0)
    x1, m = tr.hmc_step((x, beta), eps=0.02, nleapfrog=3)
    assert x1.shape == (4, 4 * 256 * 9) and torch.isfinite(m['loss']) and not x1.requires_grad
    x2, m = tr.eval_step((x1, beta))
    assert torch.isfinite(m['acc']).all() and 'mc_states' not in m
    before = {n: p.detach().clone() for n, p in tr.dynamics.named_parameters() if p.requires_grad}
    x3 = x2
    for _ in range(2):
        x3, m = tr.train_step((x3, beta))
        assert torch.isfinite(m['loss'])
    changed = [n for n, p in tr.dynamics.named_parameters() if p.requires_grad and not torch.equal(p, before[n])]
    assert any(n.endswith('veps.0') for n in changed) and any('vnet' in n for n in changed)
    assert all(torch.isfinite(p).all() for p in tr.dynamics.parameters())
    a, mx = lat.g.checkSU(tr._x(x3))
    assert float(mx.max()) < 1e-10, 'compat_proj at the top of every step puts the links back on SU(3)'


def test_su3_train_step_under_bf16_autocast(default_dtype):
    """BASELINE cfg 5: bf16 vnet (autocast) + fp64 lattice"""
    default_dtype(torch.float32)
    torch.manual_seed(1)
    np.random.seed(1)
    tr, lat = _su3_trainer(autocast=torch.bfloat16)
    x = lat.random().to(torch.complex128)
    xo, m = tr.train_step((x, torch.tensor(6.0)))
    assert torch.isfinite(m['loss']) and torch.isfinite(m['acc']).all()
    g = [p.grad for p in tr.dynamics.parameters() if p.grad is not None]
    assert g and all(torch.isfinite(t).all() for t in g)


def test_u1_trainer_steps(default_dtype):
    default_dtype(torch.float32)
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, LossConfig, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    from l2hmc_b200.trainers.pytorch.trainer import Trainer
    torch.manual_seed(2)
    np.random.seed(2)
    nb, shape = 64, [16, 16]
    cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=4, eps=0.1, eps_hmc=0.125, verbose=False)
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=[16, 16], activation_fn='leaky_relu', dropout_prob=0.2,
                                                      use_batch_norm=True), conv_config=None, net_weights=None)
    lat = LatticeU1(nb, shape)
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    tr = Trainer(dyn, LossConfig(use_mixed_loss=True, charge_weight=0.01), lr=1e-3, clip_val=10.0)
    x = lat.random()
    beta = torch.tensor(4.0)
    for _ in range(10):                       # thermalise with plain HMC (trainer.warmup)
        x, m = tr.hmc_step((x, beta), eps=0.1, nleapfrog=8)
    losses = []
    for _ in range(3):
        x, m = tr.train_step((x, beta))
        losses.append(float(m['loss']))
    assert all(np.isfinite(losses))
    x, m = tr.eval_step((x, beta))
    assert m['acc'].shape == (nb,) and float(x.abs().max()) <= np.pi + 1e-5


def test_trainer_warmup_thermalises_u1_towards_the_exact_plaquette():
    """Trainer.warmup (trainer.py:1699-1744): accept / reject HMC from a hot start; the U(1) plaquette approaches
    I1(beta) / I0(beta); verbose step metrics carry the lattice observables (trainer.py:922-924)"""
    from l2hmc_b200.configs import DynamicsConfig, LossConfig
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1, plaq_exact
    from l2hmc_b200.trainers.pytorch.trainer import Trainer
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        torch.manual_seed(4)
        np.random.seed(4)
        nb, shape, beta = 256, [16, 16], 2.0
        cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=8, eps=0.1, eps_hmc=0.125, verbose=True)
        lat = LatticeU1(nb, shape)
        dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
        tr = Trainer(dyn, LossConfig())
        x0 = lat.random()
        p0 = float(lat.plaqs(x0).mean())
        x = tr.warmup(beta, nsteps=60, x=x0)
        assert x.shape[0] == nb
        want = float(plaq_exact(torch.tensor(beta)))
        p1 = float(lat.plaqs(x.reshape(x0.shape)).mean())
        assert abs(p0) < 0.05 and abs(p1 - want) < 0.02, (p0, p1, want)
        assert tr.warmup(beta, nsteps=1, x=x0, nchains=7).shape[0] == 7
        _, m = tr.hmc_step((x, torch.tensor(beta)))
        assert {'plaqs', 'intQ', 'sinQ', 'dQint', 'dQsin'} <= set(m)
    finally:
        torch.set_default_dtype(old)

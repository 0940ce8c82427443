"""GPU tier: `Dynamics.reuse_force = 'always'` (the default) computes the force and the vnet inputs once per
distinct link configuration instead of once per v-update.  The forward sweep must be bit-identical to
`reuse_force = 'never'` (the reference's recompute-everything schedule, dynamics.py:1187-1228), the gradients
equal to rounding, and the number of force evaluations per fb sweep must drop from 4 nlf to 2 nlf + 1
(host logic: tests/test_reuse_force_logic.py)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def dev(a):
    return torch.from_numpy(np.ascontiguousarray(a)).to(DEV)


def _su3_dynamics(gl, units=8):
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, NetWeights, NetWeight, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    shape, nb, nlf = [int(s) for s in gl['shape']], 2, int(gl['nlf'])
    cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=nlf, eps=0.05, eps_hmc=0.1,
                         verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
    torch.manual_seed(1)
    np.random.seed(1)
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=[units], activation_fn='tanh', dropout_prob=0.0,
                                                      use_batch_norm=False),
                         conv_config=None, net_weights=NetWeights(x=NetWeight(0., 1., 1.), v=NetWeight(1., 1., 1.)))
    lat = LatticeSU3(nb, shape)
    return Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac), lat, nlf


def _count_forces(dyn):
    n = {'force': 0}
    orig = dyn.grad_potential

    def counting(x, beta):
        n['force'] += 1
        return orig(x, beta)
    dyn.grad_potential = counting
    return n


@pytest.mark.parametrize('autocast', [False, True])
def test_su3_fb_sweep_is_bit_identical_and_evaluates_fewer_forces(golden_dir, autocast):
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    old = torch.get_default_dtype()
    # bf16 autocast casts float32 modules (the reference trains its float32 nets under autocast, trainer.py:211-219);
    # the lattice stays complex128 either way
    torch.set_default_dtype(torch.float32 if autocast else torch.float64)
    try:
        gl = np.load(golden_dir / 'su3_l2hmc_f64.npz')
        dyn, lat, nlf = _su3_dynamics(gl, units=16 if autocast else 8)
        dyn.eval()
        if autocast:
            dyn.planar_sweep = 'never'        # the planar sweep has its own cache; cover the boundary-layout path first
        st = State(dev(gl['x']), dev(gl['v']), torch.tensor(float(gl['beta'])))
        outs = {}
        counts = _count_forces(dyn)
        for mode in ('never', 'always'):
            dyn.reuse_force = mode
            counts['force'] = 0
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16, enabled=autocast):
                sp, met = dyn.transition_kernel_fb(st)
            outs[mode] = (sp.x.clone(), sp.v.clone(), met['acc'].clone(), met['sumlogdet'].clone(), counts['force'])
        assert outs['never'][4] == 4 * nlf and outs['always'][4] == 2 * nlf + 1
        for a, b in zip(outs['never'][:4], outs['always'][:4]):
            assert torch.equal(a, b)
        if autocast:                          # and the planar sweep (force evaluated by ops, not grad_potential)
            dyn.planar_sweep = 'auto'
            res = {}
            for mode, pair in (('never', 'never'), ('always', 'never'), ('always', 'auto')):
                dyn.reuse_force, dyn.pair_updates = mode, pair
                with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
                    sp, met = dyn.transition_kernel_fb(st)
                res[mode, pair] = (sp.x.clone(), sp.v.clone(), met['acc'].clone(), met['sumlogdet'].clone())
            for a, b in zip(res['never', 'never'], res['always', 'never']):
                assert torch.equal(a, b)
            # paired updates (two momentum updates / both masked link updates in one pass): links and momenta keep their
            # bits; the two log-Jacobian terms of a pair are added per element in the kernel's fp32 epilogue before the
            # per-chain sum, so sumlogdet (and acc through it) moves at fp32 rounding level (the heads are bf16 GEMMs)
            ref, got = res['always', 'never'], res['always', 'auto']
            assert torch.equal(ref[0], got[0]) and torch.equal(ref[1], got[1])
            assert float((ref[3] - got[3]).abs().max()) <= 1e-6 * max(1.0, float(ref[3].abs().max()))
            assert float((ref[2] - got[2]).abs().max()) <= 1e-6
    finally:
        torch.set_default_dtype(old)


def test_su3_gradients_with_reuse_equal_default(golden_dir):
    from l2hmc_b200.dynamics.pytorch.dynamics import State
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        gl = np.load(golden_dir / 'su3_l2hmc_f64.npz')
        dyn, lat, nlf = _su3_dynamics(gl)
        dyn.train()
        grads = {}
        for mode in ('never', 'always'):
            dyn.reuse_force = mode
            dyn.zero_grad(set_to_none=True)
            x = dev(gl['x']).requires_grad_(True)
            st = State(x, dev(gl['v']), torch.tensor(float(gl['beta'])))
            sp, met = dyn.transition_kernel_fb(st)
            loss = (met['acc'] * (sp.x.flatten(1).real ** 2).sum(1)).sum() + met['sumlogdet'].sum()
            loss.backward()
            grads[mode] = ({k: p.grad.clone() for k, p in dyn.named_parameters() if p.grad is not None},
                           x.grad.clone(), float(loss))
        assert grads['never'][2] == grads['always'][2]
        assert set(grads['never'][0]) == set(grads['always'][0]) and len(grads['never'][0]) > 8
        for k, g0 in grads['never'][0].items():
            g1 = grads['always'][0][k]
            assert float((g0 - g1).abs().max()) <= 1e-10 * max(1e-30, float(g0.abs().max())), k
        assert float((grads['never'][1] - grads['always'][1]).abs().max()) <= 1e-10 * float(grads['never'][1].abs().max())
    finally:
        torch.set_default_dtype(old)

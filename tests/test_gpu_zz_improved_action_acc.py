"""GPU tier: plain HMC with an improved-action `potential_fn` (c1 != 0).  The reference integrates with the
force of the Dynamics' own c1 = 0 lattice and accepts with the energies of `potential_fn`
(dynamics.py:134,1489-1499); the golden case has acc = (8.7e-4, 0.93), so the energies matter.
(File name sorts last on purpose: it was added after the round's last GPU run.)"""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.mark.parametrize('kernel', [True, False])
def test_hmc_accepts_with_the_energies_of_potential_fn(golden_dir, kernel):
    from l2hmc_b200.configs import DynamicsConfig
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        g = np.load(golden_dir / 'su3_c1_f64.npz')
        shape, nb, beta, c1 = [int(s) for s in g['shape']], g['x'].shape[0], float(g['beta']), float(g['c1'])
        lat = LatticeSU3(nb, shape, c1=c1)
        lat.rect_kernel = kernel
        cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=2, eps=0.05, eps_hmc=0.05,
                             verbose=False, use_split_xnets=False, use_separate_networks=False)
        dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
        x0 = torch.from_numpy(g['hmc2_x0']).to(DEV)
        v0 = torch.from_numpy(g['hmc2_v0']).to(DEV)
        st = State(x0, v0, torch.tensor(beta))
        with torch.no_grad():
            sp, met = dyn.transition_kernel_hmc(st, eps=0.01, nleapfrog=3)
            h0 = dyn.hamiltonian(st)
        xp = sp.x.cpu().numpy().reshape(g['hmc2_x0'].shape)
        vp = sp.v.cpu().numpy().reshape(g['hmc2_x0'].shape)
        assert np.abs(xp - g['hmc2_x'].reshape(xp.shape)).max() < 1e-12
        assert np.abs(vp - g['hmc2_v'].reshape(vp.shape)).max() < 1e-12
        assert np.allclose(h0.cpu().numpy(), g['hmc2_h0'], rtol=1e-12)
        assert np.allclose(met['acc'].cpu().numpy(), g['hmc2_acc'], rtol=1e-7)     # |H| ~ 4e3 at 1e-12 relative
    finally:
        torch.set_default_dtype(old)

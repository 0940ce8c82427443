"""Shared builders for the GPU tier and the multi-GPU check scripts (imported as
`tests._helpers`; `tests/` is a package, the repo root is on sys.path via conftest)."""


def _su3_trainer(nb=4, shape=(4, 4, 4, 4), units=(16,), nlf=2, autocast=None):
    from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, NetWeights, NetWeight, LossConfig, get_input_spec
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    from l2hmc_b200.trainers.pytorch.trainer import Trainer
    cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=list(shape), nleapfrog=nlf, eps=0.05, eps_hmc=0.05,
                         verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=list(units), activation_fn='tanh', dropout_prob=0.0,
                                                      use_batch_norm=False),
                         conv_config=None, net_weights=NetWeights(x=NetWeight(0., 1., 1.), v=NetWeight(1., 1., 1.)),
                         build_unused_su3_xnet=False)
    lat = LatticeSU3(nb, list(shape))
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    tr = Trainer(dyn, LossConfig(use_mixed_loss=False, charge_weight=0.0, rmse_weight=0.1, plaq_weight=0.1), lr=1e-3,
                 autocast_dtype=autocast)
    return tr, lat

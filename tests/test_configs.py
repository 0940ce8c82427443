"""CPU tier: boundary dataclasses behave like the reference's configs.py for the
YAML nodes of conf/dynamics, conf/network, conf/conv, conf/net_weights."""
import yaml

from l2hmc_b200 import configs as c


def test_su3_yaml_nodes_instantiate():
    # conf/dynamics/su3.yaml, conf/network/su3.yaml, conf/net_weights/su3.yaml (values from SURVEY section 5)
    dyn = c.from_target_dict(yaml.safe_load('''
_target_: l2hmc.configs.DynamicsConfig
group: SU3
latvolume: [4, 4, 4, 4]
nchains: 8
nleapfrog: 2
eps: 0.01
use_split_xnets: false
use_separate_networks: false
merge_directions: true
'''))
    assert dyn.xshape == (8, 4, 4, 4, 4, 4, 3, 3) and dyn.xdim == 4 * 256 * 9 and dyn.vshape[-1] == 8
    net = c.from_target_dict({'_target_': 'l2hmc.configs.NetworkConfig', 'units': [256], 'activation_fn': 'tanh',
                              'dropout_prob': 0.0, 'use_batch_norm': False})
    assert list(net.units) == [256]
    nw = c.from_target_dict({'_target_': 'l2hmc.configs.NetWeights', 'x': {'s': 0.0, 't': 1.0, 'q': 1.0},
                             'v': {'s': 1.0, 't': 1.0, 'q': 1.0}})
    assert nw.x.s == 0.0 and nw.v.q == 1.0
    spec = c.get_input_spec(dyn)
    assert spec.vnet['x'] == [4 * 256 * 8]


def test_u1_defaults():
    dyn = c.DynamicsConfig(nchains=128, group='U1', latvolume=[16, 16], nleapfrog=8, eps=0.1, eps_hmc=None)
    assert dyn.xshape == (128, 2, 16, 16) and dyn.xdim == 512 and dyn.eps_hmc == 1.0 / 8
    conv = c.ConvolutionConfig(filters=[8, 16], sizes=None, pool=None)
    assert list(conv.sizes) == [2, 2] and list(conv.pool) == [2, 2]
    spec = c.get_input_spec(dyn)
    assert spec.xnet == {'x': [512, 2], 'v': [512]}

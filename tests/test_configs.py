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


def test_to_str_fragments():
    """configs.py:133-139,294-305,416-448,531-548: the strings the reference builds its output directories from"""
    assert c.NetworkConfig(units=[16, 16], activation_fn='relu', dropout_prob=0.2).to_str() == 'net-16-16_dp-0.2_bn-True'
    assert c.ConvolutionConfig(filters=[8, 16], sizes=[5, 3], pool=[2, 2]).to_str() == 'conv-8-16_5-3_2-2'
    assert c.ConvolutionConfig().to_str() == 'conv-None' and c.ConvolutionConfig(filters=[]).to_str() == ''
    assert c.LossConfig(use_mixed_loss=True, charge_weight=0.01).to_str() == 'qw-0.0_pw-0.0_rw-0.0_aw-0.0_mixed-True'
    assert c.InputSpec(xshape=(8, 2, 16, 16)).to_str() == '8-2-16-16'
    assert c.NetWeights(x=c.NetWeight(0., 1., 1.), v=c.NetWeight(1., 1., 1.)).to_str() == 'nwx-s0.0t1.0q1.0-nwv-s1.0t1.0q1.0'
    assert c.list_to_str([0.5, 1.0]) == '0.5-1.0' and c.list_to_str(['a', 'b']) == 'a-b'
    assert c.NetWeight(1., 2., 3.).to_dict() == {'s': 1., 't': 2., 'q': 3.}
    assert '"units": [4]' in c.NetworkConfig(units=[4], activation_fn='relu', dropout_prob=0.0).to_json()

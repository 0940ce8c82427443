"""CPU tier: the host-side mirror of the reference's network module tree
(l2hmc_b200/network/pytorch/network.py <-> network/pytorch/network.py:151-551).

The reference's own `state_dict` (frozen in tests/golden by oracle/make_golden.py) must load
into our `LeapfrogLayer` with no missing or unexpected key and identical shapes (SURVEY
section 8 f-3, appendix B trap 15), and the module's eval-mode forward must agree with the
numpy oracle (oracle/network.py) on the same weights.  Only the module tree is exercised
here -- it is library plumbing (torch.nn); the public factory still refuses to run without a
CUDA device, which the last test pins."""
import numpy as np
import pytest
import torch

from l2hmc_b200 import configs as c
from l2hmc_b200.network.pytorch import network as net
from oracle import network as onet

CONV = dict(filters=[4, 8, 8], sizes=[3, 2, 2], pool=[2, 2, 2])


def _layer(gu, name, key, dtype):
    """our LeapfrogLayer for one net of the U(1) golden config (make_golden.py: units [16, 12],
    leaky_relu, dropout 0.2, batch norm; `conv` adds the 3-stage periodic conv stack)"""
    shape = [int(s) for s in gu['shape']]
    xshape = (3, 2, *shape)
    xdim = 2 * shape[0] * shape[1]
    is_x = key.startswith('xnet')
    ncfg = c.NetworkConfig(units=[16, 12], activation_fn='leaky_relu', dropout_prob=0.2, use_batch_norm=True)
    ccfg = c.ConvolutionConfig(**CONV) if name == 'conv' else None
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        layer = net.LeapfrogLayer(xshape=xshape, network_config=ncfg,
                                  input_shapes={'x': [xdim, 2], 'v': [xdim]} if is_x else {'x': [xdim], 'v': [xdim]},
                                  net_weight=c.NetWeight(1., 1., 1.), conv_config=ccfg)
        layer.eval()
        with torch.no_grad():      # materialise the lazy layers, as the reference's dummy call does (network.py:572-631)
            layer((torch.zeros(2, 4 if is_x else 2, *shape), torch.zeros(2, xdim)))
    finally:
        torch.set_default_dtype(old)
    return layer, xshape, xdim, is_x


@pytest.mark.parametrize('tag', ['f32', 'f64'])
@pytest.mark.parametrize('name', ['dense', 'conv'])
def test_reference_state_dict_loads_and_forward_matches_oracle(golden_dir, tag, name):
    gu = np.load(golden_dir / f'u1_{tag}.npz')
    dtype = torch.float64 if tag == 'f64' else torch.float32
    pre = f'{name}/sd/'
    sd_all = {k[len(pre):]: gu[k] for k in gu.files if k.startswith(pre)}
    rng = np.random.default_rng(11)
    for key in ('xnet.0.first', 'xnet.1.second', 'vnet.0', 'vnet.1'):
        sd = onet.sub_state_dict(sd_all, key)
        layer, xshape, xdim, is_x = _layer(gu, name, key, dtype)
        ours = layer.state_dict()
        assert set(ours) == set(sd), sorted(set(ours) ^ set(sd))
        for k, w in sd.items():
            assert tuple(ours[k].shape) == tuple(w.shape), k
            assert ours[k].dtype == torch.from_numpy(np.asarray(w)).dtype, k
        res = layer.load_state_dict({k: torch.from_numpy(np.asarray(w)) for k, w in sd.items()}, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        npdt = np.float64 if tag == 'f64' else np.float32
        x = rng.uniform(-1, 1, (3, 4 if is_x else 2, *xshape[2:])).astype(npdt)
        v = rng.standard_normal((3, xdim)).astype(npdt)
        with torch.no_grad():
            s, t, q = layer((torch.from_numpy(x), torch.from_numpy(v)))
        so, to, qo = onet.leapfrog_layer(x, v, sd, activation='leaky_relu', use_batch_norm=True,
                                         conv=CONV if name == 'conv' else None,
                                         conv_in_shape=(4 if is_x else 2, *xshape[2:]))
        tol = 1e-12 if tag == 'f64' else 2e-5
        for a, b in ((s, so), (t, to), (q, qo)):
            assert a.shape == (3, xdim)
            assert float(np.max(np.abs(a.numpy() - b))) <= tol * max(1.0, float(np.abs(b).max()))


def test_su3_vnet_state_dict_keys(golden_dir):
    """SU(3) golden config (units [8], tanh, no batch norm): our vnet has exactly the reference's parameters"""
    gl = np.load(golden_dir / 'su3_l2hmc_f64.npz')
    shape = [int(s) for s in gl['shape']]
    V = int(np.prod(shape))
    want = {k[len('grad/vnet.'):] for k in gl.files if k.startswith('grad/vnet.')}
    ncfg = c.NetworkConfig(units=[8], activation_fn='tanh', dropout_prob=0.0, use_batch_norm=False)
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float64)
    try:
        layer = net.LeapfrogLayer(xshape=(2, 4, *shape, 3, 3), network_config=ncfg,
                                  input_shapes={'x': [4 * V * 8], 'v': [4 * V * 8]}, net_weight=c.NetWeight(1., 1., 1.))
        with torch.no_grad():
            layer((torch.zeros(2, 4, *shape, 8), torch.zeros(2, 4, *shape, 8)))
    finally:
        torch.set_default_dtype(old)
    assert {k for k, _ in layer.named_parameters()} == want
    for k, p in layer.named_parameters():
        assert tuple(p.shape) == tuple(gl[f'grad/vnet.{k}'].shape), k
    assert layer.xdim == 4 * V * 9 and layer.scale.layer.weight.shape == (4 * V * 9, 8)


def test_periodic_padding_and_scaled_tanh_match_oracle():
    rng = np.random.default_rng(5)
    x = rng.standard_normal((2, 3, 5, 4))
    for size in (1, 2, 4):
        got = net.PeriodicPadding(size)(torch.from_numpy(x)).numpy()
        assert np.array_equal(got, onet.periodic_pad(x, size))
    st = net.ScaledTanh(6, 10).double()
    with torch.no_grad():
        st.coeff.copy_(torch.from_numpy(rng.standard_normal(tuple(st.coeff.shape))))
    z = rng.standard_normal((4, 6))
    want = np.exp(st.coeff.detach().numpy()) * np.tanh(z @ st.layer.weight.detach().numpy().T
                                                       + st.layer.bias.detach().numpy())
    assert np.allclose(st(torch.from_numpy(z)).detach().numpy(), want, rtol=1e-13, atol=1e-14)


def test_unknown_activation_and_factory_without_gpu():
    with pytest.raises(ValueError):
        net.activation_fn('gelu2')
    if torch.cuda.is_available():
        pytest.skip('GPU present: the factory builds')
    dyn = c.DynamicsConfig(nchains=2, group='U1', latvolume=[4, 4], nleapfrog=1, eps=0.1)
    fac = net.NetworkFactory(input_spec=c.get_input_spec(dyn),
                             network_config=c.NetworkConfig(units=[4], activation_fn='relu', dropout_prob=0.0,
                                                            use_batch_norm=False))
    with pytest.raises(RuntimeError, match='no CPU fallback'):
        fac.build_networks(1, False, group=None)


def test_module_level_helpers():
    """network.py:49-148,209-237,554-569 and dynamics.py:77-99, lattice.py:33-38: the small free functions"""
    from l2hmc_b200.dynamics.pytorch import dynamics as d
    from l2hmc_b200.lattice.su3.pytorch import lattice as ls
    lin = torch.nn.Linear(3, 2)
    seq = torch.nn.Sequential(lin, torch.nn.Sequential(torch.nn.Tanh()))
    tree = net.nested_children(seq)
    assert set(tree) == {'0', '1'} and tree['0'] == {'Linear': lin} and list(tree['1']['0']) == ['Tanh']
    net.zero_weights(lin)
    assert float(lin.weight.detach().abs().max()) == 0 and float(lin.bias.detach().abs().max()) == 0
    net.init_weights(lin, 'xavier_normal')
    assert float(lin.weight.detach().abs().max()) > 0
    net.init_weights(lin, 'zeros')
    assert float(lin.weight.detach().abs().max()) == 0
    net.init_all(seq, torch.nn.init.constant_, 0.5)
    assert all(bool((p == 0.5).all()) for p in seq.parameters())
    net.init_all_by_shape(seq, {'default': lambda p: torch.nn.init.constant_(p, 1.0),
                                '2': lambda p: torch.nn.init.constant_(p, 2.0)})
    assert bool((lin.weight == 2.0).all()) and bool((lin.bias == 1.0).all())
    assert net.calc_output_size((16, 16), 5) == (12, 12) and net.calc_output_size((24, 24), 2, stride=2) == (12, 12)
    assert net.calc_output_size((15, 9), (3, 2), stride=2, pad=1) == (8, 5)
    x = torch.tensor([[0.0, np.pi / 2]])
    assert torch.allclose(net.xy_repr(x), torch.tensor([[[1.0, 0.0], [0.0, 1.0]]]), atol=1e-7)
    layer = net.get_network((2, 2, 4, 4), c.NetworkConfig(units=[4], activation_fn='relu', dropout_prob=0.0,
                                                          use_batch_norm=False))
    assert isinstance(layer, net.LeapfrogLayer) and layer.xdim == 32
    a = d.random_angle((100,), requires_grad=True)
    assert a.requires_grad and float(a.abs().max()) <= np.pi
    assert torch.allclose(d.to_u1(torch.tensor([3 * np.pi / 2, -3 * np.pi / 2, 0.3])),
                          torch.tensor([-np.pi / 2, np.pi / 2, 0.3]), atol=1e-6)
    assert ls.pbc((4, -1, 2, 7), (4, 4, 4, 4)) == [0, 3, 2, 3]
    m = np.array([[1 + 2j, 3], [4j, 5]])
    assert np.array_equal(ls.mat_adj(m), m.conj().T)

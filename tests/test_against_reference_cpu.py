"""CPU tier, only where the reference copy `oracle/_ref` exists (this container; built by
`__graft_entry__.build()` from /root/reference): the boundary dataclasses are compared field by field with
the reference's own `configs.py` for a sweep of constructor arguments, and the pure-torch module helpers
with the reference's functions.  Skipped elsewhere (nothing at run time reads /root/reference)."""
import itertools

import numpy as np
import pytest
import torch

from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason='oracle/_ref not built here')


@pytest.fixture(scope='module')
def ref():
    old = torch.get_default_dtype()
    r = ref_shim.load_reference(torch.float64)
    yield r
    torch.set_default_dtype(old)


def test_dynamics_config_matches_reference(ref):
    from l2hmc_b200 import configs as c
    cases = [dict(group='U1', latvolume=[16, 16]), dict(group='U1', latvolume=[8, 6]),
             dict(group='SU3', latvolume=[4, 4, 4, 8]), dict(group='SU3', latvolume=[2, 3, 4, 5])]
    opts = itertools.product([4, 7], [1, 3], [0.05, 0.2], [None, 0.3], [True, False], [True, False])
    for (nchains, nlf, eps, eps_hmc, merge, split), base in itertools.product(opts, cases):
        kw = dict(nchains=nchains, nleapfrog=nlf, eps=eps, eps_hmc=eps_hmc, merge_directions=merge,
                  use_split_xnets=split, use_separate_networks=split, verbose=False, **base)
        a, b = c.DynamicsConfig(**kw), ref.DynamicsConfig(**kw)
        for f in ('nchains', 'group', 'nleapfrog', 'eps', 'eps_hmc', 'use_ncp', 'verbose', 'eps_fixed', 'use_split_xnets',
                  'use_separate_networks', 'merge_directions', 'xdim', 'dim'):
            assert getattr(a, f) == getattr(b, f), (f, kw)
        assert tuple(a.xshape) == tuple(b.xshape) and list(a.latvolume) == list(b.latvolume)
        assert a.to_str() == b.to_str(), kw


def test_network_and_loss_configs_match_reference(ref):
    from l2hmc_b200 import configs as c
    for units, act, dp, bn in (([16, 16], 'relu', 0.2, True), ([256], 'tanh', 0.0, False)):
        a = c.NetworkConfig(units=units, activation_fn=act, dropout_prob=dp, use_batch_norm=bn)
        b = ref.NetworkConfig(units=units, activation_fn=act, dropout_prob=dp, use_batch_norm=bn)
        assert a.to_str() == b.to_str() and a.asdict() == b.asdict()
    for kw in (dict(filters=[8, 16, 32], sizes=[5, 3, 3], pool=[2, 2, 2]), dict(filters=[4, 8]), dict()):
        a, b = c.ConvolutionConfig(**kw), ref.ConvolutionConfig(**kw)
        assert a.to_str() == b.to_str()
        assert (a.sizes is None and b.sizes is None) or list(a.sizes) == list(b.sizes)
        assert (a.pool is None and b.pool is None) or list(a.pool) == list(b.pool)
    for kw in (dict(), dict(use_mixed_loss=True, charge_weight=0.0, rmse_weight=0.1, plaq_weight=0.1)):
        assert c.LossConfig(**kw).to_str() == ref.cfgs.LossConfig(**kw).to_str()
    nw = dict(x=dict(s=0.0, t=1.0, q=1.0), v=dict(s=1.0, t=0.5, q=2.0))
    assert c.NetWeights(**nw).to_str() == ref.NetWeights(**nw).to_str()
    assert c.NetWeights(**nw).to_dict() == ref.NetWeights(**nw).to_dict()
    for xshape in ((8, 2, 16, 16), (4, 4, 2, 2, 2, 2, 3, 3)):
        a, b = c.InputSpec(xshape=xshape), ref.InputSpec(xshape=xshape)
        assert a.to_str() == b.to_str() and a.xdim == b.xdim
        assert tuple(a.vshape) == tuple(b.vshape) and a.vdim == b.vdim


def test_periodic_padding_and_conv_stack_shapes_match_reference(ref):
    """the module tree under one random state_dict: our ConvStack / LeapfrogLayer output == the reference's"""
    import importlib
    from l2hmc_b200 import configs as c
    from l2hmc_b200.network.pytorch import network as net
    rnet = importlib.import_module('l2hmc.network.pytorch.network')
    xshape, xdim = (3, 2, 8, 6), 96
    conv = dict(filters=[4, 8, 8], sizes=[3, 2, 2], pool=[2, 2, 2])
    ncfg = dict(units=[16, 12], activation_fn='leaky_relu', dropout_prob=0.2, use_batch_norm=True)
    torch.manual_seed(0)
    theirs = rnet.LeapfrogLayer(xshape=xshape, network_config=ref.NetworkConfig(**ncfg),
                                input_shapes={'x': [xdim, 2], 'v': [xdim]}, conv_config=ref.ConvolutionConfig(**conv))
    ours = net.LeapfrogLayer(xshape=xshape, network_config=c.NetworkConfig(**ncfg),
                             input_shapes={'x': [xdim, 2], 'v': [xdim]}, conv_config=c.ConvolutionConfig(**conv))
    x, v = torch.randn(3, 4, 8, 6), torch.randn(3, xdim)
    theirs.eval()
    ours.eval()
    with torch.no_grad():
        want = theirs((x, v))
        ours((x, v))                                  # materialise lazy layers
        res = ours.load_state_dict(theirs.state_dict(), strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        got = ours((x, v))
    for a, b in zip(got, want):
        assert torch.allclose(a, b.to(a.dtype), rtol=1e-12, atol=1e-12)
    pad = rnet.PeriodicPadding(2)
    assert torch.equal(net.PeriodicPadding(2)(x), pad(x))

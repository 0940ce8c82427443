"""CPU tier: the HOST wiring of the tensor-core dense / conv layers -- autograd.TCDense (bf16 and fp32-x3 modes, ragged
input pairs, the deferred weight gradients), ConvPeriodic / PoolAct, ConvStack's walk over the reference's layer
list, ops.gemm_f32's segment lists -- on CPU stand-ins of the kernels (tests/cpu_emulation.py), against the same
modules on plain torch.  The kernels themselves are covered by tests/test_gpu_{gemm,dense,conv}.py."""
import numpy as np
import pytest
import torch

from tests import cpu_emulation as emu


def _su3_like_net(units, act, seed=0):
    from l2hmc_b200.configs import NetworkConfig
    from l2hmc_b200.network.pytorch.network import LeapfrogLayer
    torch.manual_seed(seed)
    nb, shape = 4, (2, 2, 2, 3)
    net = LeapfrogLayer((nb, 4, *shape, 3, 3), NetworkConfig(units=list(units), activation_fn=act, dropout_prob=0.0,
                                                             use_batch_norm=False))
    with torch.no_grad():
        _ = net((torch.zeros(2, 4, *shape, 8), torch.zeros(2, 4, *shape, 8)))
        net.scale.coeff.normal_(0, 0.1)
        net.transf.coeff.normal_(0, 0.1)
    return net, nb, shape


@pytest.mark.parametrize('act', ['tanh', 'swish', 'leaky_relu'])
@pytest.mark.parametrize('defer', [False, True])
def test_tcdense_fp32_mode_equals_the_torch_modules(monkeypatch, act, defer):
    from l2hmc_b200 import autograd as ag
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        with emu.tensor_core_layers_on_cpu(monkeypatch):
            net, nb, shape = _su3_like_net((20, 12), act, seed=3)
            g = torch.Generator().manual_seed(4)
            x0, f0 = torch.randn(nb, 4, *shape, 8, generator=g), torch.randn(nb, 4, *shape, 8, generator=g)
            ws = [torch.randn(nb, net.xdim, generator=g) for _ in range(3)]
            res = {}
            for mode in ('never', 'auto'):
                net.tc_dense = mode
                net.zero_grad(set_to_none=True)
                x, f = x0.clone().requires_grad_(True), f0.clone().requires_grad_(True)
                assert net.tensor_core_dense(x, f) == (None if mode == 'never' else 'x3')
                s, t, q = net((x, f))
                loss = (s * ws[0]).sum() + (t * ws[1]).sum() + (q * ws[2]).sum()
                monkeypatch.setattr(ag, 'DEFER_HEAD_GRADS', defer and mode == 'auto')
                loss.backward()
                monkeypatch.setattr(ag, 'DEFER_HEAD_GRADS', False)
                assert not ag._PENDING_DENSE
                res[mode] = ([o.detach() for o in (s, t, q)], x.grad, f.grad,
                             {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None})
            a, b = res['auto'], res['never']
            for o1, o0 in zip(a[0], b[0]):
                assert float((o1 - o0).abs().max()) <= 2e-5 * max(1e-3, float(o0.abs().max()))
            for k in (1, 2):
                assert float((a[k] - b[k]).abs().max()) <= 5e-5 * max(1e-6, float(b[k].abs().max()))
            assert set(a[3]) == set(b[3]) and len(a[3]) >= 10
            for n_, g0 in b[3].items():
                assert float((a[3][n_] - g0).abs().max()) <= 5e-5 * max(1e-6, float(g0.abs().max())), n_
    finally:
        torch.set_default_dtype(old)


def test_tcdense_bf16_mode_under_autocast(monkeypatch):
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        with emu.tensor_core_layers_on_cpu(monkeypatch):
            monkeypatch.setattr(torch, 'is_autocast_enabled', lambda *a: True)
            monkeypatch.setattr(torch, 'get_autocast_dtype', lambda *a: torch.bfloat16)
            net, nb, shape = _su3_like_net((32,), 'tanh', seed=5)
            g = torch.Generator().manual_seed(6)
            x = torch.randn(nb, 4, *shape, 8, generator=g).to(torch.bfloat16).requires_grad_(True)
            f = torch.randn(nb, 4, *shape, 8, generator=g).to(torch.bfloat16).requires_grad_(True)
            assert net.tensor_core_dense(x, f) == 'bf16'
            s, t, q = net((x, f))
            (s.float().sum() + t.float().sum() + q.float().sum()).backward()
            got = {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None}
            # the same computation spelled out in float32 on the bf16-rounded operands
            il = net.input_layer
            bf = lambda t_: t_.detach().to(torch.bfloat16).float()  # noqa: E731
            z = torch.tanh(bf(x).reshape(nb, -1) @ bf(il.xlayer.weight).t() + bf(f).reshape(nb, -1) @ bf(il.vlayer.weight).t()
                           + il.xlayer.bias + il.vlayer.bias).to(torch.bfloat16).float()
            s_ref = net.scale.coeff.exp() * torch.tanh(z @ bf(net.scale.layer.weight).t() + net.scale.layer.bias)
            assert float((s.float() - s_ref).abs().max()) <= 2e-2 * max(1e-3, float(s_ref.abs().max()))
            assert len(got) >= 10 and all(torch.isfinite(v).all() for v in got.values())
    finally:
        torch.set_default_dtype(old)


@pytest.mark.parametrize('cfg', [((8, 16, 32, 64, 128), (5, 3, 3, 3, 2), (2, 2, 2, 2, 2), 16), ((4, 8, 6), (3, 2, 2), (2, 2, 2), 8)])
@pytest.mark.parametrize('act', ['leaky_relu', 'swish'])
def test_conv_stack_walk_equals_the_torch_modules(monkeypatch, cfg, act):
    from l2hmc_b200.configs import ConvolutionConfig
    from l2hmc_b200.network.pytorch.network import ConvStack, activation_fn
    filters, sizes, pool, L = cfg
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        with emu.tensor_core_layers_on_cpu(monkeypatch):
            torch.manual_seed(7)
            cs = ConvStack((3, 2, L, L), ConvolutionConfig(filters=list(filters), sizes=list(sizes), pool=list(pool)),
                           activation_fn(act))
            cs.conv_precision = 'fp32'
            with torch.no_grad():
                _ = cs(torch.zeros(2, 4, L, L))
            g = torch.Generator().manual_seed(8)
            nb = 3
            x0 = torch.randn(nb, 4, L, L, generator=g)
            wv = torch.randn(nb, 2 * L * L, generator=g)
            res = {}
            for mode in ('never', 'auto'):
                cs.tc_conv = mode
                cs.zero_grad(set_to_none=True)
                x = x0.clone().requires_grad_(True)
                assert cs.tensor_core_mode(x) == (None if mode == 'never' else 'x3')
                y = cs(x)
                (y * wv).sum().backward()
                res[mode] = (y.detach(), x.grad, {n_: p.grad.clone() for n_, p in cs.named_parameters()})
            a, b = res['auto'], res['never']
            assert a[0].shape == b[0].shape
            assert float((a[0] - b[0]).abs().max()) <= 3e-5 * max(1e-3, float(b[0].abs().max()))
            assert float((a[1] - b[1]).abs().max()) <= 1e-4 * max(1e-6, float(b[1].abs().max()))
            for n_, g0 in b[2].items():
                assert float((a[2][n_] - g0).abs().max()) <= 1e-4 * max(1e-6, float(g0.abs().max())), n_
    finally:
        torch.set_default_dtype(old)


def test_gemm_f32_segment_lists(monkeypatch):
    """ops.gemm_f32: six (three) products per pair from the bf16x3 (x2) planes, chunks of five pairs accumulate"""
    from l2hmc_b200 import ops
    with emu.tensor_core_layers_on_cpu(monkeypatch):
        g = torch.Generator().manual_seed(9)
        a = [torch.randn(32, 24, generator=g) for _ in range(7)]      # [K, M]
        b = [torch.randn(32, 40, generator=g) for _ in range(7)]      # [K, N]
        want = sum(x.double().t() @ y.double() for x, y in zip(a, b))
        got = ops.gemm_f32([ops.split_bf16x3(t) for t in a], [ops.split_bf16x3(t) for t in b], False, False)
        assert float((got.double() - want).abs().max()) <= 4e-6 * float(want.abs().max())
        got2 = ops.gemm_f32([ops.split_bf16x3(t)[:2] for t in a], [ops.split_bf16x3(t)[:2] for t in b], False, False)
        e2 = float((got2.double() - want).abs().max() / want.abs().max())
        assert 1e-7 < e2 <= 2e-4
        with pytest.raises(ops.L2BError):
            ops.gemm_f32([ops.split_bf16x3(t) for t in a], [ops.split_bf16x3(t) for t in b], False, False, act='tanh')

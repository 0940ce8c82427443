"""GPU tier: the U(1) xnet's convolution stack on the hand-written path (periodic-padding gather + tensor-core GEMM +
pooling kernels, csrc/l2b_conv.cu, autograd.ConvPeriodic / PoolAct) against the same module on torch / cuDNN
(reference network/pytorch/network.py:151-172, 240-346): values and every gradient, fp32 nets (bf16x3 GEMM, fp32
accuracy) and bf16 autocast; the building blocks against torch restatements of their definitions."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


@pytest.fixture(autouse=True)
def _exact_fp32_reference():
    """the torch reference of these tests must be real fp32 (cuDNN would otherwise use TF32 on this GPU)"""
    old = torch.backends.cudnn.allow_tf32
    torch.backends.cudnn.allow_tf32 = False
    yield
    torch.backends.cudnn.allow_tf32 = old


def test_im2col_is_periodic_padding_plus_unfold():
    from l2hmc_b200 import ops
    from l2hmc_b200.network.pytorch.network import PeriodicPadding
    g = torch.Generator(device='cpu').manual_seed(2)
    for (nb, C, H, W, n) in [(3, 4, 6, 5, 3), (2, 5, 4, 4, 5), (2, 8, 7, 3, 2), (1, 3, 5, 6, 4)]:
        x = torch.randn(nb, C, H, W, generator=g).to(DEV)
        want = torch.nn.functional.unfold(PeriodicPadding(n - 1)(x), n)            # [nb, C n n, OH OW]
        want = want.transpose(1, 2).reshape(nb * (H + n - 1) * (W + n - 1), C * n * n)
        for nchw in (True, False):
            xin = x if nchw else x.permute(0, 2, 3, 1).contiguous()
            c3 = ops.conv_im2col(xin, n, nchw, 3)
            assert c3.shape[2] % 8 == 0 and float(c3[:, :, C * n * n:].abs().sum()) == 0.0
            got = c3.double().sum(0)[:, :C * n * n]
            assert float((got - want.double()).abs().max()) <= 2.0 ** -22 * float(want.abs().max())
            c1 = ops.conv_im2col(xin.to(torch.bfloat16), n, nchw, 1)
            assert torch.equal(c1[0][:, :C * n * n], want.to(torch.bfloat16))
        # adjoint: <col2im(d), x> == <d, im2col(x)>
        d = torch.randn(want.shape[0], C * n * n + 3, generator=g).to(DEV)
        gx = ops.conv_col2im(d, x, n, True)
        lhs = float((gx.double() * x.double()).sum())
        rhs = float((d[:, :C * n * n].double() * want.double()).sum())
        assert abs(lhs - rhs) <= 1e-5 * max(1.0, abs(rhs))
        gx2 = ops.conv_col2im(d, x.permute(0, 2, 3, 1).contiguous(), n, False)
        assert float((gx2.permute(0, 3, 1, 2) - gx).abs().max()) == 0.0


def test_tap_major_gather_is_a_column_permutation_of_the_channel_major_one():
    """k_order 1 (columns (kh, kw, ci), vector loads when Cin % 8 == 0 on NHWC) against k_order 0 (columns
    (ci, kh, kw)): bit-equal after the permutation, for the vector path, the scalar fallbacks (Cin % 8 != 0, NCHW,
    an unaligned view) and the adjoint"""
    from l2hmc_b200 import ops
    g = torch.Generator(device='cpu').manual_seed(4)
    for (nb, C, H, W, n) in [(3, 8, 6, 5, 3), (2, 16, 4, 4, 2), (2, 32, 12, 12, 3), (2, 5, 4, 6, 3), (1, 24, 7, 3, 4)]:
        K = C * n * n
        x = torch.randn(nb, H, W, C, generator=g).to(DEV)
        for planes, xin in ((3, x), (2, x), (1, x.to(torch.bfloat16))):
            c0 = ops.conv_im2col(xin, n, False, planes)
            c1 = ops.conv_im2col(xin, n, False, planes, tap_major=True)
            want = c0[:, :, :K].reshape(planes, -1, C, n * n).transpose(2, 3).reshape(planes, -1, K)
            assert torch.equal(c1[:, :, :K], want)
            assert float(c1[:, :, K:].abs().sum()) == 0.0
        # NCHW input and an unaligned NHWC view take the scalar path of the same kernel
        xn = x.permute(0, 3, 1, 2).contiguous()
        c0 = ops.conv_im2col(xn, n, True, 3)
        c1 = ops.conv_im2col(xn, n, True, 3, tap_major=True)
        assert torch.equal(c1[:, :, :K], c0[:, :, :K].reshape(3, -1, C, n * n).transpose(2, 3).reshape(3, -1, K))
        big = torch.randn(nb * H * W * C + 1, generator=g).to(DEV)
        xv = big[1:].view(nb, H, W, C)
        assert torch.equal(ops.conv_im2col(xv, n, False, 3, tap_major=True), ops.conv_im2col(xv.clone(), n, False, 3, tap_major=True))
        # adjoint
        M = nb * (H + n - 1) * (W + n - 1)
        d = torch.randn(M, K + 3, generator=g).to(DEV)
        d_tap = torch.cat([d[:, :K].reshape(M, C, n * n).transpose(1, 2).reshape(M, K), d[:, K:]], 1).contiguous()
        for dt in (torch.float32, torch.bfloat16):
            g0 = ops.conv_col2im(d.to(dt), x, n, False)
            g1 = ops.conv_col2im(d_tap.to(dt), x, n, False, tap_major=True)
            assert torch.equal(g0, g1)


@pytest.mark.parametrize('act', [None, 'tanh', 'relu', 'swish', 'leaky_relu', 'elu'])
def test_pool_act_matches_torch(act):
    from l2hmc_b200 import autograd as ag
    fn = {None: lambda t: t, 'tanh': torch.tanh, 'relu': torch.relu, 'swish': torch.nn.functional.silu,
          'leaky_relu': lambda t: torch.nn.functional.leaky_relu(t, 0.01), 'elu': torch.nn.functional.elu}[act]
    g = torch.Generator(device='cpu').manual_seed(3)
    for (nb, H, W, C, p) in [(3, 6, 6, 5, 2), (2, 7, 5, 4, 2), (2, 9, 9, 3, 3)]:
        x0 = torch.randn(nb, H, W, C, generator=g).to(DEV)
        w = torch.randn(nb, H // p, W // p, C, generator=g).to(DEV)
        x = x0.clone().requires_grad_(True)
        y = ag.PoolAct.apply(x, p, act)
        (y * w).sum().backward()
        xr = x0.clone().requires_grad_(True)
        yr = fn(torch.nn.functional.max_pool2d(xr.permute(0, 3, 1, 2), p)).permute(0, 2, 3, 1)
        (yr * w).sum().backward()
        assert float((y - yr).abs().max()) <= 1e-6
        assert float((x.grad - xr.grad).abs().max()) <= 1e-6


def _stack(filters, sizes, pool, act, T=8, X=8, seed=0):
    from l2hmc_b200.configs import ConvolutionConfig
    from l2hmc_b200.network.pytorch.network import ConvStack, activation_fn
    torch.manual_seed(seed)
    cs = ConvStack((4, 2, T, X), ConvolutionConfig(filters=list(filters), sizes=list(sizes), pool=list(pool)),
                   activation_fn(act)).to(DEV)
    with torch.no_grad():
        _ = cs(torch.zeros(2, 4, T, X, device=DEV))
    return cs


@pytest.mark.parametrize('act', ['leaky_relu', 'tanh', 'swish'])
@pytest.mark.parametrize('cfg', [((8, 16, 32, 64, 128), (5, 3, 3, 3, 2), (2, 2, 2, 2, 2), 16),     # conf/conv/default.yaml
                                 ((4, 8), (3, 2), (2, 2), 8), ((6, 4, 12), (2, 3, 2), (2, 2, 2), 8)])
def test_conv_stack_fp32_matches_torch(act, cfg):
    filters, sizes, pool, L = cfg
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        cs = _stack(filters, sizes, pool, act, T=L, X=L, seed=5)
        cs.conv_precision = 'fp32'                       # the exact mode (the default is the TF32-class bf16x2 mode)
        g = torch.Generator(device='cpu').manual_seed(6)
        nb = 5
        x0 = torch.randn(nb, 4, L, L, generator=g).to(DEV)
        wv = torch.randn(nb, 2 * L * L, generator=g).to(DEV)
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
        assert a[0].shape == b[0].shape == (nb, 2 * L * L) and a[0].dtype == torch.float32
        assert float((a[0] - b[0]).abs().max()) <= 3e-5 * max(1e-3, float(b[0].abs().max()))
        assert float((a[1] - b[1]).abs().max()) <= 1e-4 * max(1e-6, float(b[1].abs().max()))
        assert set(a[2]) == set(b[2])
        for n_, g0 in b[2].items():
            assert float((a[2][n_] - g0).abs().max()) <= 1e-4 * max(1e-6, float(g0.abs().max())), n_
    finally:
        torch.set_default_dtype(old)


def test_conv_stack_bf16_autocast_is_as_close_to_fp32_as_torch_autocast():
    """under bf16 autocast two implementations differ by bf16 noise amplified through six layers (a max-pool tap can
    flip): both are measured against the fp32 stack; ours must not be further from it than torch's own bf16 path"""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        L = 16
        cs = _stack((8, 16, 32, 64, 128), (5, 3, 3, 3, 2), (2, 2, 2, 2, 2), 'leaky_relu', T=L, X=L, seed=7)
        g = torch.Generator(device='cpu').manual_seed(8)
        nb = 6
        x0 = torch.randn(nb, 4, L, L, generator=g).to(DEV)
        wv = torch.randn(nb, 2 * L * L, generator=g).to(DEV)
        res = {}
        for name, mode, cast in (('fp32', 'never', False), ('torch', 'never', True), ('ours', 'auto', True)):
            cs.tc_conv = mode
            cs.zero_grad(set_to_none=True)
            x = x0.clone().requires_grad_(True)
            with torch.autocast('cuda', dtype=torch.bfloat16, enabled=cast):
                if name == 'ours':
                    assert cs.tensor_core_mode(x) == 'bf16'
                y = cs(x)
            (y.float() * wv).sum().backward()
            res[name] = (y.detach().float(), x.grad.float(), {n_: p.grad.float().clone() for n_, p in cs.named_parameters()})

        def err(a, b):
            return float((a - b).norm() / b.norm().clamp(min=1e-12))
        ref, ours, lib = res['fp32'], res['ours'], res['torch']
        assert err(ours[0], ref[0]) <= 2.0 * err(lib[0], ref[0]) + 1e-3
        assert err(ours[1], ref[1]) <= 2.0 * err(lib[1], ref[1]) + 1e-3
        for n_, g0 in ref[2].items():
            assert err(ours[2][n_], g0) <= 2.0 * err(lib[2][n_], g0) + 1e-3, n_
    finally:
        torch.set_default_dtype(old)


def test_conv_stack_tf32_class_mode():
    """the default conv_precision = 'tf32' (bf16x2 operands, 16 mantissa bits): within TF32-class distance of the fp32 stack,
    closer to it than cuDNN's own TF32 convolution (what the reference runs on this GPU by default)"""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        L = 16
        cs = _stack((8, 16, 32, 64, 128), (5, 3, 3, 3, 2), (2, 2, 2, 2, 2), 'leaky_relu', T=L, X=L, seed=9)
        g = torch.Generator(device='cpu').manual_seed(10)
        x0 = torch.randn(6, 4, L, L, generator=g).to(DEV)
        cs.tc_conv = 'never'
        with torch.no_grad():
            ref = cs(x0)
            torch.backends.cudnn.allow_tf32 = True
            lib = cs(x0)
            torch.backends.cudnn.allow_tf32 = False
            cs.tc_conv = 'auto'                           # default conv_precision = 'tf32'
            got = cs(x0)
        e_got = float((got - ref).abs().max() / ref.abs().max())
        e_lib = float((lib - ref).abs().max() / ref.abs().max())
        assert e_got <= 2e-4 and e_got <= max(e_lib, 1e-5) * 1.5, (e_got, e_lib)
    finally:
        torch.set_default_dtype(old)


def test_u1_train_step_launches_no_library_gemm_or_conv():
    """U(1) L2HMC training with the reference's default conv stack in fp32 (its default precision): every GEMM and
    every convolution of a training step is one of ours (kernel-name census with torch.profiler)"""
    from torch.profiler import ProfilerActivity, profile
    from l2hmc_b200.configs import (ConvolutionConfig, DynamicsConfig, LossConfig, NetworkConfig, get_input_spec)
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    from l2hmc_b200.trainers.pytorch.trainer import Trainer
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        torch.manual_seed(3)
        np.random.seed(3)
        nb, shape = 16, [16, 16]
        cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=2, eps=0.1, eps_hmc=None, use_ncp=True,
                             verbose=False, eps_fixed=False, use_split_xnets=True, merge_directions=True,
                             use_separate_networks=True)
        fac = NetworkFactory(input_spec=get_input_spec(cfg),
                             network_config=NetworkConfig(units=[16, 16], activation_fn='leaky_relu', dropout_prob=0.0,
                                                          use_batch_norm=False),
                             conv_config=ConvolutionConfig(filters=[8, 16, 32, 64, 128], sizes=[5, 3, 3, 3, 2],
                                                           pool=[2, 2, 2, 2, 2]), net_weights=None)
        lat = LatticeU1(nb, shape)
        dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
        tr = Trainer(dyn, LossConfig(use_mixed_loss=True, charge_weight=0.01), lr=1e-3, clip_val=1.0)
        x, beta = lat.random(), torch.tensor(4.0)
        for _ in range(2):
            xo, m = tr.train_step((x, beta))
        assert torch.isfinite(m['loss'])
        torch.cuda.synchronize()
        try:
            with profile(activities=[ProfilerActivity.CUDA]) as prof:
                tr.train_step((x, beta))
                torch.cuda.synchronize()
            names = [e.key for e in prof.key_averages() if e.key]
        except Exception as e:
            pytest.skip(f'torch.profiler unavailable: {e}')
        if not names:
            pytest.skip('profiler returned no kernels')
        markers = ('nvjet', 'cutlass', 'gemm', 'cublas', 'cudnn', 'conv', 'xmma', 'implicit', 'sm90_', 'sm100_', 'wgrad', 'dgrad')
        lib = [n for n in names if any(k in n.lower() for k in markers) and 'l2b::' not in n]
        ours = [n for n in names if 'l2b::' in n]
        assert any('k_im2col_periodic' in n for n in ours) and any('k_gemm_bf16' in n for n in ours)
        assert any('k_col2im_periodic' in n for n in ours) and any('k_pool_act' in n for n in ours)
        assert not lib, lib
    finally:
        torch.set_default_dtype(old)


@pytest.mark.parametrize('act', ['leaky_relu', 'tanh'])
def test_xnet_with_conv_stack_matches_the_numpy_oracle(act):
    """a whole U(1) xnet LeapfrogLayer with the default conv stack in fp32 -- conv blocks, pooling, the input pair,
    hidden Linears and the three heads, all on the hand-written kernels -- against oracle/network.py (float64 numpy
    restatement of network.py:240-346, 454-551) on the same weights"""
    from l2hmc_b200.configs import ConvolutionConfig, NetworkConfig
    from l2hmc_b200.network.pytorch.network import LeapfrogLayer
    from oracle import network as onet
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        torch.manual_seed(21)
        T = X = 16
        nb = 4
        conv = dict(filters=[8, 16, 32, 64, 128], sizes=[5, 3, 3, 3, 2], pool=[2, 2, 2, 2, 2])
        net = LeapfrogLayer((nb, 2, T, X), NetworkConfig(units=[16, 16], activation_fn=act, dropout_prob=0.0,
                                                         use_batch_norm=False),
                            conv_config=ConvolutionConfig(**conv)).to(DEV)
        net.eval()
        with torch.no_grad():
            _ = net((torch.zeros(2, 4, T, X, device=DEV), torch.zeros(2, 2 * T * X, device=DEV)))
            net.scale.coeff.normal_(0, 0.1)
            net.transf.coeff.normal_(0, 0.1)
        g = torch.Generator(device='cpu').manual_seed(22)
        x = torch.randn(nb, 4, T, X, generator=g)
        v = torch.randn(nb, 2 * T * X, generator=g)
        assert net.tensor_core_dense(x.to(DEV), v.to(DEV)) == 'x3'
        assert net.input_layer.conv_stack.tensor_core_mode(x.to(DEV)) == 'x3'
        net.input_layer.conv_stack.conv_precision = 'fp32'
        with torch.no_grad():
            s, t, q = net((x.to(DEV), v.to(DEV)))
        sd = {k: p.detach().double().cpu().numpy() for k, p in net.state_dict().items()}
        ws, wt, wq = onet.leapfrog_layer(x.double().numpy(), v.double().numpy(), sd, activation=act, conv=conv,
                                         conv_in_shape=[4, T, X])
        for got, want in ((s, ws), (t, wt), (q, wq)):
            assert float(np.abs(got.double().cpu().numpy() - want).max()) <= 5e-5 * max(1e-2, float(np.abs(want).max()))
    finally:
        torch.set_default_dtype(old)

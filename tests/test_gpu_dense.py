"""GPU tier: the dense layers of the networks on the hand-written tensor-core GEMM (`autograd.TCDense`,
csrc/l2b_gemm.cu) against torch's own bf16 autocast path -- what the reference runs for BASELINE cfg 5
(network/pytorch/network.py:415-451, 536-548 under trainer.py:211-219 autocast) -- values and every gradient;
and a kernel-name census of an L2HMC train / eval step: no library GEMM (cuBLAS `nvjet` / `cutlass` / `gemm`) left."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _net(xshape, units, act='tanh', seed=0):
    from l2hmc_b200.configs import NetworkConfig
    from l2hmc_b200.network.pytorch.network import LeapfrogLayer
    torch.manual_seed(seed)
    net = LeapfrogLayer(xshape, NetworkConfig(units=list(units), activation_fn=act, dropout_prob=0.0,
                                              use_batch_norm=False)).to(DEV)
    nb = 2
    with torch.no_grad():
        _ = net((torch.zeros((nb, *xshape[1:6], 8), device=DEV), torch.zeros((nb, *xshape[1:6], 8), device=DEV)))
        net.scale.coeff.normal_(0, 0.1)
        net.transf.coeff.normal_(0, 0.1)
    return net


@pytest.mark.parametrize('act', ['tanh', 'relu', 'swish', 'leaky_relu', 'elu'])
@pytest.mark.parametrize('units', [(64,), (20, 12)])
def test_tcdense_fp32_network_matches_torch_fp32(act, units):
    """fp32 nets without autocast (the reference's default precision): TCDense in 'x3' mode against torch's fp32
    Linear + autograd, values and every gradient at fp32 accuracy"""
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        nb, shape = 6, (2, 4, 2, 3)
        xshape = (nb, 4, *shape, 3, 3)
        net = _net(xshape, units, act=act, seed=7)
        g = torch.Generator(device='cpu').manual_seed(8)
        x0 = torch.randn(nb, 4, *shape, 8, generator=g).to(DEV)
        f0 = torch.randn(nb, 4, *shape, 8, generator=g).to(DEV)
        ws = [torch.randn(nb, net.xdim, generator=g).to(DEV) for _ in range(3)]
        res = {}
        for mode in ('never', 'auto'):
            net.tc_dense = mode
            net.zero_grad(set_to_none=True)
            x, f = x0.clone().requires_grad_(True), f0.clone().requires_grad_(True)
            assert net.tensor_core_dense(x, f) == (None if mode == 'never' else 'x3')
            s, t, q = net((x, f))
            loss = (s * ws[0]).sum() + (t * ws[1]).sum() + (q * ws[2]).sum()
            loss.backward()
            res[mode] = ([o.detach() for o in (s, t, q)], x.grad, f.grad,
                         {n: p.grad.clone() for n, p in net.named_parameters() if p.grad is not None})
        a, b = res['auto'], res['never']
        for o1, o0 in zip(a[0], b[0]):
            assert o1.dtype == torch.float32 and float((o1 - o0).abs().max()) <= 2e-5 * max(1e-3, float(o0.abs().max()))
        for k in (1, 2):
            assert float((a[k] - b[k]).abs().max()) <= 5e-5 * max(1e-6, float(b[k].abs().max()))
        assert set(a[3]) == set(b[3]) and len(a[3]) >= 10
        for n_, g0 in b[3].items():
            assert float((a[3][n_] - g0).abs().max()) <= 5e-5 * max(1e-6, float(g0.abs().max())), n_
    finally:
        torch.set_default_dtype(old)


@pytest.mark.parametrize('act', ['tanh', 'relu', 'swish', 'leaky_relu', 'elu'])
@pytest.mark.parametrize('units', [(64,), (48, 32)])
def test_tcdense_network_matches_torch_autocast(act, units):
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        nb, shape = 6, (2, 4, 2, 3)
        xshape = (nb, 4, *shape, 3, 3)
        net = _net(xshape, units, act=act, seed=3)
        g = torch.Generator(device='cpu').manual_seed(4)
        x0 = torch.randn(nb, 4, *shape, 8, generator=g).to(DEV).to(torch.bfloat16)
        f0 = torch.randn(nb, 4, *shape, 8, generator=g).to(DEV).to(torch.bfloat16)
        ws = [torch.randn(nb, net.xdim, generator=g).to(DEV) * 0.01 for _ in range(3)]
        res = {}
        for mode in ('never', 'auto'):
            net.tc_dense = mode
            net.zero_grad(set_to_none=True)
            x, f = x0.clone().requires_grad_(True), f0.clone().requires_grad_(True)
            with torch.autocast('cuda', dtype=torch.bfloat16):
                s, t, q = net((x, f))
            loss = (s.float() * ws[0]).sum() + (t.float() * ws[1]).sum() + (q.float() * ws[2]).sum()
            loss.backward()
            res[mode] = ([o.detach().float() for o in (s, t, q)], x.grad.float(), f.grad.float(),
                         {n: p.grad.float().clone() for n, p in net.named_parameters() if p.grad is not None})
        a, b = res['auto'], res['never']
        for o1, o0 in zip(a[0], b[0]):
            assert float((o1 - o0).abs().max()) <= 3e-2 * max(1e-3, float(o0.abs().max()))
        for k in (1, 2):
            assert float((a[k] - b[k]).abs().max()) <= 5e-2 * max(1e-6, float(b[k].abs().max()))
        assert set(a[3]) == set(b[3]) and len(a[3]) >= 10
        for n_, g0 in b[3].items():
            g1 = a[3][n_]
            assert g1.shape == g0.shape
            assert float((g1 - g0).abs().max()) <= 5e-2 * max(1e-6, float(g0.abs().max())), n_
    finally:
        torch.set_default_dtype(old)


def _kernel_names(fn, nwarm=2):
    from torch.profiler import ProfilerActivity, profile
    for _ in range(nwarm):
        fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn()
        torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages() if getattr(e, 'device_type', None) is not None]
    return [n for n in names if n]


LIBRARY_GEMM_MARKERS = ('nvjet', 'cutlass', 'gemm', 'cublas', 'cudnn', 'sm90_', 'sm100_', 'xmma')


@pytest.mark.parametrize('step', ['eval', 'train'])
@pytest.mark.parametrize('units', [(32,), (32, 16)])
def test_l2hmc_step_launches_no_library_gemm(step, units):
    """BASELINE cfg 5 (SU(3), bf16 nets): every GEMM of a train / eval step is one of ours"""
    from tests._helpers import _su3_trainer
    old = torch.get_default_dtype()
    torch.set_default_dtype(torch.float32)
    try:
        torch.manual_seed(1)
        np.random.seed(1)
        tr, lat = _su3_trainer(nb=8, units=units, autocast=torch.bfloat16)
        x = lat.random().to(torch.complex128)
        beta = torch.tensor(6.0)
        def eval_fn():                   # the reference's eval_step has no autocast of its own: the caller wraps it
            with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
                return tr.eval_step((x, beta))
        fn = (lambda: tr.train_step((x, beta))) if step == 'train' else eval_fn
        try:
            names = _kernel_names(fn)
        except Exception as e:                                   # no CUPTI on this box
            pytest.skip(f'torch.profiler unavailable: {e}')
        if not names:
            pytest.skip('profiler returned no kernels')
        ours = [n for n in names if 'l2b::' in n]
        lib = [n for n in names if any(m in n.lower() for m in LIBRARY_GEMM_MARKERS) and 'l2b::' not in n]
        assert any('k_gemm_bf16' in n or 'k_su3_input_gemm' in n for n in ours), ours
        assert any('k_heads_vupdate' in n for n in ours)
        assert not lib, lib
    finally:
        torch.set_default_dtype(old)

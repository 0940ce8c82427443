"""GPU tier: the general bf16 tensor-core GEMM (l2b_gemm_bf16) that carries the dense layers of the networks and their
backward passes (network/pytorch/network.py:415-422, 489-493, 538-548 under ATen autograd in the reference).
Checked against a float64 evaluation of the SAME bf16-rounded operands (the kernel accumulates in fp32), for every
operand orientation (K-major / MN-major), segments, split-K, ragged extents, bias / activation / accumulate."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
DEV = 'cuda:0'


def _mk(shape, seed):
    g = torch.Generator(device='cpu').manual_seed(seed)
    return torch.randn(shape, generator=g).to(DEV).to(torch.bfloat16)


def _ref(a_list, b_list, a_k, b_k, bias=None, act=None):
    acc = None
    for a, b in zip(a_list, b_list):
        A = a.double() if a_k else a.double().t()
        B = b.double() if b_k else b.double().t()
        d = A @ B.t()
        acc = d if acc is None else acc + d
    if bias is not None:
        acc = acc + bias.double()
    if act == 'tanh':
        acc = torch.tanh(acc)
    elif act == 'relu':
        acc = torch.relu(acc)
    elif act == 'swish':
        acc = acc * torch.sigmoid(acc)
    elif act == 'leaky_relu':
        acc = torch.nn.functional.leaky_relu(acc, 0.01)
    elif act == 'elu':
        acc = torch.nn.functional.elu(acc)
    return acc


CASES = [
    # M, N, K, a_kmajor, b_kmajor, nseg, splits
    (32, 256, 512, True, True, 1, 1),        # Linear forward
    (32, 256, 4096, True, True, 2, 0),       # two input Linears, automatic split-K
    (32, 256, 1024, True, False, 3, 0),      # dz = sum_heads dY W  (W contracted over its rows)
    (256, 256, 64, False, False, 1, 1),      # dW = dY^T X
    (384, 512, 96, False, False, 1, 1),      # several M and N tiles, K tail
    (40, 72, 200, True, True, 1, 1),         # ragged M, N (multiple of 8), K tail
    (136, 328, 136, True, False, 1, 3),      # ragged + forced split-K
    (130, 40, 24, False, True, 1, 1),        # MN-major A, K-major B
    (2, 16, 16, True, True, 1, 1),           # tiny
    (5, 12, 20, True, False, 1, 1),          # extents that need zero padding on the Python side
    (7, 9, 3, False, False, 1, 1),
    (256, 320, 32, False, False, 16, 1),     # dW over 16 updates of 32 chains: two segments share one 64-row stage
    (136, 72, 24, False, False, 5, 1),       # ... odd segment count, K < 32, ragged
    (200, 64, 32, False, False, 7, 2),       # ... with split-K
    (32, 1024, 256, True, False, 1, 1),      # skinny M, short K: computed as D^T (dX of the input layer)
    (16, 520, 72, True, True, 2, 1),
    (64, 200, 40, False, False, 1, 1),
]


@pytest.mark.parametrize('M,N,K,a_k,b_k,nseg,splits', CASES)
def test_gemm_matches_float64_of_the_same_bf16_operands(M, N, K, a_k, b_k, nseg, splits):
    from l2hmc_b200 import ops
    a = [_mk((M, K) if a_k else (K, M), 10 + s) for s in range(nseg)]
    b = [_mk((N, K) if b_k else (K, N), 20 + s) for s in range(nseg)]
    want = _ref(a, b, a_k, b_k)
    got = ops.gemm_bf16(a, b, a_k, b_k, out_dtype=torch.float32, splits=splits)
    assert got.shape == (M, N) and got.dtype == torch.float32
    scale = float(want.abs().max())
    assert float((got.double() - want).abs().max()) <= 2e-6 * scale * max(1.0, np.sqrt(K * nseg) / 8), (M, N, K)
    gb = ops.gemm_bf16(a, b, a_k, b_k, out_dtype=torch.bfloat16, splits=splits)
    assert gb.dtype == torch.bfloat16
    assert float((gb.double() - want).abs().max()) <= 1e-2 * scale


@pytest.mark.parametrize('act', [None, 'tanh', 'relu', 'swish', 'leaky_relu', 'elu'])
@pytest.mark.parametrize('splits', [1, 4])
@pytest.mark.parametrize('M', [48, 160])
def test_gemm_bias_activation(act, splits, M):
    from l2hmc_b200 import ops
    N, K = 64, 1024
    a, b = _mk((M, K), 1) * 0.05, _mk((N, K), 2)
    bias = torch.linspace(-1, 1, N, device=DEV)
    want = _ref([a], [b], True, True, bias=bias, act=act)
    got = ops.gemm_bf16(a, b, True, True, out_dtype=torch.float32, bias=bias, act=act, splits=splits)
    assert float((got.double() - want).abs().max()) <= 5e-6 * max(1.0, float(want.abs().max()))


@pytest.mark.parametrize('odt', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('splits', [1, 2])
def test_gemm_accumulates_into_a_strided_view(odt, splits):
    from l2hmc_b200 import ops
    M, N, K = 256, 128, 64
    gy, x = _mk((K, M), 3), _mk((K, N), 4)
    buf = torch.randn(M, N + 64, device=DEV).to(odt)
    view = buf[:, 32:32 + N]                       # row stride N + 64, 16-byte aligned offset
    before = buf.clone()
    ops.linear_dw(gy, x, out=view, accumulate=True)
    want = before[:, 32:32 + N].double() + gy.double().t() @ x.double()
    tol = 1e-5 if odt == torch.float32 else 2e-2
    assert float((view.double() - want).abs().max()) <= tol * float(want.abs().max())
    assert torch.equal(buf[:, :32], before[:, :32]) and torch.equal(buf[:, 32 + N:], before[:, 32 + N:])


def test_linear_triplet_equals_torch_autograd():
    """forward, dX and dW of one Linear against torch's own bf16 Linear + autograd (the reference's path)"""
    from l2hmc_b200 import ops
    nb, fin, fout = 32, 256, 512
    x = _mk((nb, fin), 5).requires_grad_(True)
    lin = torch.nn.Linear(fin, fout).to(DEV).to(torch.bfloat16)
    y = lin(x)
    gy = _mk((nb, fout), 6)
    y.backward(gy)
    y2 = ops.linear_fwd(x.detach(), lin.weight.detach(), lin.bias.detach())
    assert float((y2.float() - y.detach().float()).abs().max()) <= 2e-2 * float(y.detach().float().abs().max())
    dx = ops.linear_dx(gy, lin.weight.detach())
    dw = ops.linear_dw(gy, x.detach())
    assert float((dx.float() - x.grad.float()).abs().max()) <= 2e-2 * float(x.grad.float().abs().max())
    assert float((dw - lin.weight.grad.float()).abs().max()) <= 2e-2 * float(lin.weight.grad.float().abs().max())


def test_gemm_rejects_bad_arguments():
    from l2hmc_b200 import ops
    from l2hmc_b200._lib import L2BError
    a, b = _mk((8, 16), 1), _mk((8, 24), 2)
    with pytest.raises(L2BError):
        ops.gemm_bf16(a, b, True, True)                       # contraction lengths differ
    with pytest.raises(L2BError):
        ops.gemm_bf16(a.float(), a.float(), True, True)       # not bf16
    with pytest.raises(L2BError):
        ops.gemm_bf16(a.cpu(), a.cpu(), True, True)           # no CPU fallback


@pytest.mark.parametrize('M,N,K,a_k,b_k,npair', [
    (32, 256, 512, True, True, 1), (32, 64, 4096, True, True, 2), (40, 72, 200, True, False, 1),
    (256, 136, 32, False, False, 7), (5, 12, 20, True, True, 1), (130, 40, 24, False, True, 1),
])
def test_fp32_accurate_gemm_from_bf16x3_splits(M, N, K, a_k, b_k, npair):
    """gemm_f32: six bf16 products per operand pair reproduce an fp32 GEMM (relative error ~1e-6 of the row scale)"""
    from l2hmc_b200 import ops
    g = torch.Generator(device='cpu').manual_seed(31)
    a = [torch.randn((M, K) if a_k else (K, M), generator=g).to(DEV) for _ in range(npair)]
    b = [torch.randn((N, K) if b_k else (K, N), generator=g).to(DEV) for _ in range(npair)]
    for t in a + b:
        s3 = ops.split_bf16x3(t)
        rec = s3.double().sum(0)[:, :t.shape[1]]
        assert float((rec - t.double()).abs().max()) <= 2.0 ** -22 * float(t.abs().max())
        assert float(s3[:, :, t.shape[1]:].abs().max() if s3.shape[2] > t.shape[1] else 0.0) == 0.0
    want = sum((x.double() if a_k else x.double().t()) @ (y.double() if b_k else y.double().t()).t() for x, y in zip(a, b))
    got = ops.gemm_f32([ops.split_bf16x3(t) for t in a], [ops.split_bf16x3(t) for t in b], a_k, b_k)
    got = got[:M, :N]
    scale = float(want.abs().max())
    assert float((got.double() - want).abs().max()) <= 4e-6 * scale
    # an fp32 cuBLAS GEMM of the same operands is no closer
    lib = sum((x if a_k else x.t()) @ (y if b_k else y.t()).t() for x, y in zip(a, b))
    assert float((got.double() - want).abs().max()) <= 8 * float((lib.double() - want).abs().max()) + 1e-7 * scale


@pytest.mark.parametrize('odt', [torch.float32, torch.bfloat16])
@pytest.mark.parametrize('a_k,b_k', [(False, False), (True, True), (True, False)])
def test_skinny_gemm_computed_as_transpose_accumulates_and_adds_bias(odt, a_k, b_k):
    """M <= 64 with a short K loop runs as D^T inside the kernel: accumulate, bias and activation must still address
    out[m][n] / bias[n]"""
    from l2hmc_b200 import ops
    M, N, K = 16, 136, 72
    a = _mk((M, K) if a_k else (K, M), 41)
    b = _mk((N, K) if b_k else (K, N), 42)
    bias = torch.linspace(-0.5, 0.5, N, device=DEV)
    want = _ref([a], [b], a_k, b_k, bias=bias, act='tanh')
    got = ops.gemm_bf16(a, b, a_k, b_k, out_dtype=odt, bias=bias, act='tanh', splits=1)
    tol = 5e-6 if odt == torch.float32 else 1e-2
    assert float((got.double() - want).abs().max()) <= tol * max(1.0, float(want.abs().max()))
    base = torch.randn(M, N, device=DEV).to(odt)
    out = base.clone()
    ops.gemm_bf16(a, b, a_k, b_k, out=out, accumulate=True, splits=1)
    want2 = base.double() + _ref([a], [b], a_k, b_k)
    assert float((out.double() - want2).abs().max()) <= (1e-5 if odt == torch.float32 else 3e-2) * float(want2.abs().max())

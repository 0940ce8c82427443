"""The GEMM shapes of one SU(3) L2HMC training step (BASELINE cfg 5: 8^4, 32 chains / GPU, units [256], bf16) on
l2b_gemm_bf16 vs torch (cuBLAS), CUDA-event timed, operands rotated through > L2-sized pools.  One JSON line per shape."""
import json
import sys

import torch

sys.path.insert(0, str(__import__("pathlib").Path(__file__).resolve().parents[1]))

from l2hmc_b200 import ops

dev = torch.device('cuda', 0)
nb, H, KIN, XD, NU = 32, 256, 131072, 147456, 16
POOL = 3


def bf(*shape):
    return [torch.randn(*shape, device=dev).to(torch.bfloat16) for _ in range(POOL)]


def timed(fn, reps=12):
    """GPU time per call: the calls are captured into one CUDA graph and the replay is timed, so that host-side
    launch overhead (ctypes, tensor-map encoding, torch dispatch) does not hide the kernels' own duration"""
    for i in range(3):
        fn(i)
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        fn(0)
    torch.cuda.current_stream().wait_stream(side)
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        for i in range(reps):
            fn(i)
    graph.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    graph.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def report(name, us_ours, us_lib, nbytes):
    print(json.dumps({'shape': name, 'ours_us': round(us_ours, 1), 'cublas_us': round(us_lib, 1),
                      'algorithmic_MB': round(nbytes / 1e6, 1), 'ours_GBps': round(nbytes / us_ours / 1e3, 1),
                      'cublas_GBps': round(nbytes / us_lib / 1e3, 1)}), flush=True)


which = set(sys.argv[1:])


def want(n):
    return not which or n in which


if want('input_fwd'):
    ax, af, wx, wv = bf(nb, KIN), bf(nb, KIN), bf(H, KIN), bf(H, KIN)
    bias = torch.zeros(H, device=dev)
    t1 = timed(lambda i: ops.gemm_bf16([ax[i % POOL], af[i % POOL]], [wx[i % POOL], wv[i % POOL]], True, True,
                                       bias=bias, act='tanh'))
    t0 = timed(lambda i: torch.tanh(ax[i % POOL] @ wx[i % POOL].t() + af[i % POOL] @ wv[i % POOL].t()))
    report('input_fwd  z[32,256] = act(ax Wx^T + af Wv^T), K = 2 x 131072', t1, t0, 2 * (H + nb) * KIN * 2)
    del ax, af, wx, wv

if want('dz'):
    g, w = [bf(nb, XD) for _ in range(3)], [bf(XD, H) for _ in range(3)]
    t1 = timed(lambda i: ops.linear_dx([g[k][i % POOL] for k in range(3)], [w[k][i % POOL] for k in range(3)]))
    t0 = timed(lambda i: g[0][i % POOL] @ w[0][i % POOL] + g[1][i % POOL] @ w[1][i % POOL] + g[2][i % POOL] @ w[2][i % POOL])
    report('dz  [32,256] = sum_3 g_h[32,147456] W_h[147456,256]', t1, t0, 3 * (H + nb) * XD * 2)
    del g, w

if want('dact'):
    gz, w = bf(nb, H), bf(H, KIN)
    t1 = timed(lambda i: ops.linear_dx(gz[i % POOL], w[i % POOL]))
    t0 = timed(lambda i: gz[i % POOL] @ w[i % POOL])
    report('dact  [32,131072] = gz[32,256] W_in[256,131072]', t1, t0, (H + nb) * KIN * 2)
    del gz, w

if want('dw_heads'):
    gs = [[torch.randn(nb, XD, device=dev).to(torch.bfloat16) for _ in range(NU)] for _ in range(2)]
    zs = [[torch.randn(nb, H, device=dev).to(torch.bfloat16) for _ in range(NU)] for _ in range(2)]
    out = torch.empty(XD, H, device=dev, dtype=torch.bfloat16)
    t1 = timed(lambda i: ops.gemm_bf16(gs[i % 2], zs[i % 2], False, False, out=out))
    gc, zc = [torch.cat(g) for g in gs], [torch.cat(z) for z in zs]
    t0 = timed(lambda i: torch.mm(gc[i % 2].t(), zc[i % 2], out=out))
    t0c = timed(lambda i: torch.mm(torch.cat(gs[i % 2]).t(), torch.cat(zs[i % 2]), out=out))
    report('dW_head  [147456,256] = sum_16 g_u[32,147456]^T z_u[32,256] (cuBLAS: on a pre-concatenated stash)', t1, t0,
           (NU * nb + H) * XD * 2)
    report('dW_head  same, cuBLAS incl. the torch.cat of the stash', t1, t0c, (NU * nb + H) * XD * 2)
    del gs, zs, gc, zc, out

if want('dw_in'):
    gs = [[torch.randn(nb, H, device=dev).to(torch.bfloat16) for _ in range(NU)] for _ in range(2)]
    xs = [[torch.randn(nb, KIN, device=dev).to(torch.bfloat16) for _ in range(NU)] for _ in range(2)]
    out = torch.empty(H, KIN, device=dev, dtype=torch.float32)
    t1 = timed(lambda i: ops.gemm_bf16(gs[i % 2], xs[i % 2], False, False, out=out))
    gc, xc = [torch.cat(g) for g in gs], [torch.cat(x) for x in xs]
    t0 = timed(lambda i: torch.mm(gc[i % 2].t(), xc[i % 2]).float())
    report('dW_in  [256,131072] fp32 = sum_16 g_u[32,256]^T x_u[32,131072]', t1, t0, NU * nb * KIN * 2 + H * KIN * 4)
    del gs, xs, gc, xc, out

if want('hidden'):
    z, w = bf(256, H), bf(H, H)
    bias = torch.zeros(H, device=dev)
    t1 = timed(lambda i: ops.gemm_bf16(z[i % POOL], w[i % POOL], True, True, bias=bias, act='tanh'))
    t0 = timed(lambda i: torch.tanh(torch.nn.functional.linear(z[i % POOL], w[i % POOL], bias.to(torch.bfloat16))))
    report('hidden  [256,256] = act(z W^T + b), K = 256', t1, t0, 3 * H * H * 2)

"""Times every k_force launch variant (l2b_set_option su3_force_variant) with CUDA
events: the kernel alone (planar entry point) and the whole trajectory.
Usage: python profiles/tune_force.py [L nb] ..."""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from l2hmc_b200 import ops, _lib  # noqa: E402

dev = 'cuda:0'
cases = [(16, 64), (8, 256)]
variants = [int(a) for a in sys.argv[1:]] or list(range(32))
out = []
REF = {}
same = None
for L, nb in cases:
    shape = [L] * 4
    torch.manual_seed(0)
    x = ops.su3_project(torch.complex(torch.randn(nb, 4, *shape, 3, 3, dtype=torch.float64, device=dev),
                                      torch.randn(nb, 4, *shape, 3, 3, dtype=torch.float64, device=dev)))
    v = ops.su3_rand_momentum(nb, shape, 1, 0, dev)
    U, P = ops.su3_aos_to_soa(x), ops.su3_aos_to_soa(v)
    links = x.numel() // 9
    dims = _lib.dims4(shape)
    for var in variants:
        _lib.set_option('su3_force_variant', var)
        nws = _lib.su3_ws_bytes(nb, shape)
        ws = torch.empty(nws, dtype=torch.uint8, device=dev)
        st = ops._stream()

        def force():
            _lib.call('l2b_su3_force_kick_planar', ops._ptr(U), ops._ptr(P), 6.0, 1e-3, None, nb, dims, _lib.L2B_F64,
                      ops._ptr(ws), nws, st)
        for _ in range(3):
            force()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(10):
            force()
        b.record()
        torch.cuda.synchronize()
        ms_f = a.elapsed_time(b) / 10
        ms_t = {}
        for fuse in (0, 1):
            _lib.set_option('su3_fuse_drift', fuse)
            try:
                for _ in range(2):
                    res = ops.su3_hmc_trajectory(x, v, 6.0, 0.1, 10)
                if fuse == 1:
                    if (L, nb) not in REF:
                        REF[(L, nb)] = res
                    same = all(torch.equal(p_, q_) for p_, q_ in zip(res, REF[(L, nb)]))
                a.record()
                for _ in range(3):
                    ops.su3_hmc_trajectory(x, v, 6.0, 0.1, 10)
                b.record()
                torch.cuda.synchronize()
                ms_t[fuse] = a.elapsed_time(b) / 3
            except Exception:
                ms_t[fuse] = float('nan')
        r = dict(L=L, nb=nb, variant=var, force_ms=round(ms_f, 4), force_GBps=round(432 * links / ms_f / 1e6, 1),
                 traj_ms=round(ms_t[0], 3), traj_fused_ms=round(ms_t[1], 3),
                 glups_fused=round(links * 10 / ms_t[1] * 1e3 / 1e9, 4), same_bits_as_first=same)
        print(json.dumps(r), flush=True)
        out.append(r)

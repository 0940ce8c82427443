"""A/B of k_update_gauge_bwd's register cap (option su3_gauge_bwd_minb: 3 = 168 registers, 2 = 240) at the cfg-5
training shapes (8^4 x 32 chains, element-wise mask, eps ~ 0.01): CUDA-event time per launch, same-bits check."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from l2hmc_b200 import _lib, ops  # noqa: E402
from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3  # noqa: E402

torch.manual_seed(1)
nb, shape = 32, [8, 8, 8, 8]
lat = LatticeSU3(nb, shape)
x = lat.random().to(torch.complex128)
p = lat.random_momentum()
g = torch.randn_like(x)
mask = (torch.rand(4 * 8 ** 4 * 9, device=x.device) > 0.5).float()
ref = None
for minb in (3, 2, 3, 2):
    _lib.set_option('su3_gauge_bwd_minb', minb)
    for _ in range(3):
        out = ops.su3_update_gauge_bwd(x, p, 0.0099, mask, False, g)
    evs = []
    for _ in range(20):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = ops.su3_update_gauge_bwd(x, p, 0.0099, mask, False, g)
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    us = sorted(a.elapsed_time(b) * 1e3 for a, b in evs)
    same = True if ref is None else bool(all(torch.equal(u, v) for u, v in zip(out[:3], ref[:3])))
    ref = ref or out
    print(f'minb={minb}: median {us[len(us) // 2]:.1f} us (min {us[0]:.1f}) incl. the reduce launch; same bits as first: {same}')

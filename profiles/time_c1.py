"""Improved action (c1 != 0): the rectangle-staple kernel against the ATen-op path (18 bmm + rolls per
evaluation, autograd for the force) on one GPU.  Usage: python profiles/time_c1.py [L nb]"""
import json
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3  # noqa: E402

L, nb = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (8, 64)
torch.set_default_dtype(torch.float64)
torch.manual_seed(0)
lat = LatticeSU3(nb, [L] * 4, c1=-0.331)
x = lat.random()
beta = torch.tensor(6.0)
out = {'L': L, 'nb': nb, 'links': nb * 4 * L ** 4}
res = {}
for name, flag in (('aten', False), ('kernel', True)):
    lat.rect_kernel = flag
    for _ in range(2):
        s, f = lat.action_with_grad(x, beta)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    a.record()
    for _ in range(5):
        s, f = lat.action_with_grad(x, beta)
    b.record()
    torch.cuda.synchronize()
    out[f'{name}_ms'] = a.elapsed_time(b) / 5
    res[name] = (s, f)
out['max_force_diff'] = float((res['aten'][1] - res['kernel'][1]).abs().max())
out['max_action_rel_diff'] = float(((res['aten'][0] - res['kernel'][0]) / res['aten'][0]).abs().max())
out['speedup'] = out['aten_ms'] / out['kernel_ms']
print(json.dumps(out))

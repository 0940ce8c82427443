"""Short driver for ncu captures: N HMC trajectories of the bench workload
(SU(3) 16^4, 64 chains, N_LF 10 unless overridden).  Not a benchmark: numbers
printed under a profiler are never reported."""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from l2hmc_b200 import ops  # noqa: E402

L = int(sys.argv[1]) if len(sys.argv) > 1 else 16
nb = int(sys.argv[2]) if len(sys.argv) > 2 else 64
nlf = int(sys.argv[3]) if len(sys.argv) > 3 else 10
ntraj = int(sys.argv[4]) if len(sys.argv) > 4 else 2
torch.manual_seed(9992)
shape = [L, L, L, L]
dev = 'cuda:0'
x = ops.su3_project(torch.complex(torch.randn(nb, 4, *shape, 3, 3, dtype=torch.float64, device=dev),
                                  torch.randn(nb, 4, *shape, 3, 3, dtype=torch.float64, device=dev)))
v = ops.su3_rand_momentum(nb, shape, 9992, 0, dev)
for _ in range(ntraj):
    xo, vo, en = ops.su3_hmc_trajectory(x, v, 6.0, 1.0 / nlf, nlf)
torch.cuda.synchronize()
print('H0', (en[:, 0] + en[:, 1])[:2].tolist(), 'H1', (en[:, 2] + en[:, 3])[:2].tolist())

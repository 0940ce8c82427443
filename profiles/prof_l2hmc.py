"""Driver for the SU(3) L2HMC path (BASELINE cfg 3 secondary / cfg 5):
    python profiles/prof_l2hmc.py eval  L nb nlf units   -> Dynamics.forward (no grad)
    python profiles/prof_l2hmc.py train L nb nlf units   -> Trainer.train_step (fwd+bwd+Adam), bf16 autocast nets
Prints ms per call (CUDA events, after warm-up) and, with --table, the torch
profiler's per-kernel device-time table (shares only: profiler on).  Used for
ncu captures as well; numbers printed under a profiler are never reported."""
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from l2hmc_b200.configs import (DynamicsConfig, LossConfig, NetWeight, NetWeights, NetworkConfig,  # noqa: E402
                                get_input_spec)
from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics  # noqa: E402
from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3  # noqa: E402
from l2hmc_b200.network.pytorch.network import NetworkFactory  # noqa: E402
from l2hmc_b200.trainers.pytorch.trainer import Trainer  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith('--')]
mode = args[0] if args else 'eval'
L = int(args[1]) if len(args) > 1 else 8
nb = int(args[2]) if len(args) > 2 else (256 if mode == 'eval' else 32)
nlf = int(args[3]) if len(args) > 3 else 4
units = int(args[4]) if len(args) > 4 else 256
reps = int(args[5]) if len(args) > 5 else 3
table = '--table' in sys.argv

import os  # noqa: E402
world = int(os.environ.get('WORLD_SIZE', '1'))
rank = 0
if world > 1:          # torchrun: one rank per GPU, NN gradients all-reduced in the flat bf16 bucket (as bench.py does)
    from l2hmc_b200 import dist as l2dist
    rank, _, local = l2dist.init()
    torch.cuda.set_device(local)
torch.manual_seed(9992)
np.random.seed(9992)
torch.set_default_dtype(torch.float32)
shape = [L, L, L, L]
cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=shape, nleapfrog=nlf, eps=0.01, eps_hmc=0.01,
                     verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
fac = NetworkFactory(input_spec=get_input_spec(cfg),
                     network_config=NetworkConfig(units=[units], activation_fn='tanh', dropout_prob=0.0,
                                                  use_batch_norm=False),
                     conv_config=None, net_weights=NetWeights(x=NetWeight(0., 1., 1.), v=NetWeight(1., 1., 1.)),
                     build_unused_su3_xnet=False)
lat = LatticeSU3(nb, shape)
dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
tr = Trainer(dyn, LossConfig(use_mixed_loss=True, charge_weight=0.0, rmse_weight=0.1, plaq_weight=0.1), lr=1e-4,
             clip_val=1.0, autocast_dtype=torch.bfloat16, grad_bucket_dtype=torch.bfloat16,
             cuda_graphs='--graph' in sys.argv)
torch.manual_seed(9993 + rank)
x = lat.random().to(torch.complex128)
beta = torch.tensor(6.0)


def step():
    if mode == 'eval':
        with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
            return tr.eval_step((x, beta))
    return tr.train_step((x, beta))


for _ in range(5 if '--graph' in sys.argv else 2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
e0.record()
for _ in range(reps):
    xo, m = step()
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / reps
links = nb * 4 * L ** 4
if rank == 0:
  print(f'{mode} SU3 {L}^4 nb={nb} nlf={nlf} units={units} world={world}: {ms:.2f} ms/call (wall {1e3 * (time.perf_counter() - t0) / reps:.2f}), '
        f'{links * 2 * nlf / (ms * 1e-3):.3e} link-updates/s, acc={float(m["acc"].mean()):.3f}, '
        f'mem={torch.cuda.max_memory_allocated() / 2**30:.1f} GiB')
if table:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    if rank == 0:
        print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=40 if world > 1 else 25,
                                        max_name_column_width=70))
if world > 1:
    torch.distributed.barrier()
    torch.distributed.destroy_process_group()

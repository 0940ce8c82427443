"""Driver for the U(1) L2HMC path with the reference's DEFAULT experiment config
(conf/config.yaml: 16x16, N_LF = 8, separate + split networks, conv stack
[8,16,32,64,128], units [16,16,16,16], dropout 0.2, batch norm, fp32):
    [L2B_TC=never] python profiles/prof_u1_l2hmc.py eval|train|hmc [nb] [reps] [--graph] [--table]
Prints ms per step (CUDA events + wall).  Numbers under a profiler are never reported."""
import os
import sys
import time
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from l2hmc_b200.configs import (ConvolutionConfig, DynamicsConfig, LossConfig, NetworkConfig,  # noqa: E402
                                get_input_spec)
from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics  # noqa: E402
from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1  # noqa: E402
from l2hmc_b200.network.pytorch.network import NetworkFactory  # noqa: E402
from l2hmc_b200.trainers.pytorch.trainer import Trainer  # noqa: E402

args = [a for a in sys.argv[1:] if not a.startswith('--')]
mode = args[0] if args else 'eval'
nb = int(args[1]) if len(args) > 1 else 128
reps = int(args[2]) if len(args) > 2 else 10
graph = '--graph' in sys.argv
torch.manual_seed(9992)
np.random.seed(9992)
torch.set_default_dtype(torch.float32)
L = int(os.environ.get('L2B_U1_L', '16'))
shape = [L, L]
NOCONV = os.environ.get('L2B_U1_CONV', 'default') == 'none'
cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=8, eps=0.1, eps_hmc=None, use_ncp=True,
                     verbose=False, eps_fixed=False, use_split_xnets=True, merge_directions=True,
                     use_separate_networks=True)
fac = NetworkFactory(input_spec=get_input_spec(cfg),
                     network_config=NetworkConfig(units=[16, 16, 16, 16], activation_fn='leaky_relu', dropout_prob=0.2,
                                                  use_batch_norm=True),
                     conv_config=None if NOCONV else ConvolutionConfig(filters=[8, 16, 32, 64, 128],
                                                                       sizes=[5, 3, 3, 3, 2], pool=[2, 2, 2, 2, 2]),
                     net_weights=None)
lat = LatticeU1(nb, shape)
dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
if os.environ.get('L2B_TC', 'auto') == 'never':       # A/B: dense / conv layers through torch (cuBLAS / cuDNN)
    for m_ in dyn.modules():
        m_.tc_dense = m_.tc_conv = 'never'
if os.environ.get('L2B_CONV_PRECISION'):                # 'tf32': bf16x2 operands for the convolutions
    for m_ in dyn.modules():
        m_.conv_precision = os.environ['L2B_CONV_PRECISION']
kw = {'cuda_graphs': True} if graph else {}
tr = Trainer(dyn, LossConfig(use_mixed_loss=True, charge_weight=0.01), lr=1e-3, clip_val=1.0, **kw)
x = lat.random()
beta = torch.tensor(4.0)


def step():
    if mode == 'eval':
        return tr.eval_step((x, beta))
    if mode == 'hmc':
        return tr.hmc_step((x, beta), eps=0.1, nleapfrog=16)
    return tr.train_step((x, beta))


for _ in range(5):
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
nlf_eff = 16
print(f'{mode} U1 {L}x{L} conv={"none" if NOCONV else "default"} nb={nb} graph={graph}: {ms:.3f} ms/step (wall {1e3 * (time.perf_counter() - t0) / reps:.3f}), '
      f'{nb * 2 * L * L * nlf_eff / (ms * 1e-3):.3e} link-updates/s, acc={float(m["acc"].mean()):.3f}, loss={float(m["loss"]):.4f}')
if '--table' in sys.argv:
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        step()
        torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by='cuda_time_total', row_limit=18, max_name_column_width=60))

"""How far are the fp32 U(1) L2HMC sweeps from a float64 evaluation of the SAME networks and inputs?
unfused fp32 (cuBLAS + accurate libm kernels) vs fused fp32 (l2b_u1_input_layer + l2b_u1_heads_update, SFU
intrinsics).  64x64, 64 chains, dense nets, N_LF = 8."""
import sys
from pathlib import Path

import numpy as np
import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from l2hmc_b200.configs import DynamicsConfig, NetworkConfig, get_input_spec  # noqa: E402
from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics, State  # noqa: E402
from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1  # noqa: E402
from l2hmc_b200.network.pytorch.network import NetworkFactory  # noqa: E402

nb, shape, nlf = 64, [64, 64], 8


def build(dtype):
    torch.set_default_dtype(dtype)
    torch.manual_seed(1)
    np.random.seed(1)
    cfg = DynamicsConfig(nchains=nb, group='U1', latvolume=shape, nleapfrog=nlf, eps=0.1, use_ncp=True, verbose=False,
                         use_split_xnets=True, merge_directions=True, use_separate_networks=True)
    fac = NetworkFactory(input_spec=get_input_spec(cfg),
                         network_config=NetworkConfig(units=[16, 16, 16, 16], activation_fn='leaky_relu',
                                                      dropout_prob=0.2, use_batch_norm=True), conv_config=None,
                         net_weights=None)
    lat = LatticeU1(nb, shape)
    dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    dyn.eval()
    return dyn, lat


d32, lat32 = build(torch.float32)
d64, lat64 = build(torch.float64)
d64.load_state_dict({k: v.double() if v.is_floating_point() else v for k, v in d32.state_dict().items()})
d64.masks = [m.clone() for m in d32.masks]
torch.manual_seed(5)
x = lat32.random().float()
v = torch.randn(nb, 2 * 64 * 64, device=x.device, dtype=torch.float32)
beta = torch.tensor(4.0)
with torch.no_grad():
    ref, mref = d64.transition_kernel_fb(State(x.double(), v.double(), beta))
    out = {}
    for mode in ('never', 'auto'):
        d32.fused_u1_heads = mode
        out[mode] = d32.transition_kernel_fb(State(x, v, beta))
for mode, (st, met) in out.items():
    dx = (st.x.double().reshape(nb, -1) - ref.x.reshape(nb, -1))
    dx = torch.remainder(dx + np.pi, 2 * np.pi) - np.pi
    print(f'fp32 fused={mode:5s}: |x - x64|max {float(dx.abs().max()):.3e}  |v - v64|max '
          f'{float((st.v.double().reshape(nb, -1) - ref.v.reshape(nb, -1)).abs().max()):.3e}  |sumlogdet - ref|max '
          f'{float((met["sumlogdet"].double() - mref["sumlogdet"]).abs().max()):.3e} (|ref| ~ {float(mref["sumlogdet"].abs().mean()):.1f})  '
          f'|acc - ref|max {float((met["acc"].double() - mref["acc"]).abs().max()):.3e}')

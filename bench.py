#!/usr/bin/env python
"""bench.py -- link-updates/s of the batched leapfrog hot path (BASELINE.json).

    python bench.py --gpus N --steps K --warmup W            # ours (CUDA, libl2b)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's own CPU PyTorch path

A "step" is ONE HMC trajectory (N_LF leapfrog steps + both Hamiltonians) over one
batch of synthetic chains.  Default workload = the configuration the north-star
target is quoted on, per GPU: 4-D SU(3) 16^4, 64 chains per GPU (BASELINE cfg 4:
512 chains sharded over 8 GPUs), N_LF = 10, complex128.  Chains are independent:
ranks share nothing on the data path (weak scaling, no collective); the only
collectives are the timing barrier and the max-over-ranks of the elapsed time.

Prints ONE JSON line (rank 0).  Keys beyond the base contract:
  roofline      dominant kernel (k_force_ep: one fused leapfrog step, staples + TAH +
                kick + exp + link update) algorithmic bytes by SURVEY 8(d)'s streaming
                model (6 link-sized transfers = 864 B per link-update; the kernel itself
                moves 4 = 576 B, reported next to it) / its average CUDA-event duration,
                against MEASURED_PEAKS.json hbm_gbs
  cpu_baseline  the reference's own PyTorch path on this box's host cores, on a
                bounded sample of the same workload
  e2e           same metric through the public Dynamics.apply_transition_hmc call
                with the links in pinned HOST memory: H2D of x and D2H of x_out and the
                accept probabilities inside the timed region
  parity        chain 0 of the timed batch against the oracle (the reference itself on
                the CPU when oracle/_ref travelled), after the timed region
  thermalised   the headline workload timed from a configuration relaxed by 100 HMC
                trajectories (trainer.py:1699-1744)
  secondary     the other BASELINE configs in the same run and process group: cfg 3
                (SU(3) 8^4 x 256 HMC), cfg 2 (U(1) 64x64 x 4096 HMC), cfg 3/5 L2HMC eval
                and training steps (the training step includes the NCCL all-reduce of
                the network gradients when N > 1)
  gpu_reference the reference's own .cuda() path on the same GPU (N = 1 only)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

WORKLOADS = {
    # name: (group, lattice, chains per GPU, N_LF, dtype, beta)
    'su3_16x16x16x16_nb64_nlf10_c128': ('SU3', [16, 16, 16, 16], 64, 10, 'f64', 6.0),
    'su3_8x8x8x8_nb256_nlf10_c128': ('SU3', [8, 8, 8, 8], 256, 10, 'f64', 6.0),
    'u1_64x64_nb4096_nlf10_f32': ('U1', [64, 64], 4096, 10, 'f32', 4.0),
    'u1_16x16_nb128_nlf8_f32': ('U1', [16, 16], 128, 8, 'f32', 4.0),
}
# L2HMC workloads (BASELINE cfg 3 secondary / cfg 5): a "step" is one Trainer.eval_step
# (Dynamics.forward, 2*N_LF leapfrog layers) or one Trainer.train_step (forward + loss +
# backward + gradient all-reduce over the ranks + Adam); bf16 vnet (autocast), fp64 lattice,
# network.units = [256] (conf/network/su3.yaml), N_LF = 4 (conf/su3test.yaml).
L2HMC_WORKLOADS = {
    # name: (mode, lattice, chains per GPU, N_LF, hidden units, beta)
    'su3_8x8x8x8_nb256_l2hmc_eval_bf16': ('eval', [8, 8, 8, 8], 256, 4, 256, 6.0),
    'su3_8x8x8x8_nb32_l2hmc_train_bf16': ('train', [8, 8, 8, 8], 32, 4, 256, 6.0),
}
# BASELINE cfg 1: the reference's own default experiment (conf/config.yaml: U(1) 16x16, 128 chains, N_LF = 8, fp32,
# separate + split networks, conv stack [8,16,32,64,128] / sizes [5,3,3,3,2] / pool 2, units [16,16,16,16], dropout
# 0.2, batch norm) -- the one config the CPU arm runs at full size, chain for chain.
U1_L2HMC_WORKLOADS = {
    # name: (mode, lattice, chains per GPU, N_LF, beta)
    'u1_16x16_nb128_l2hmc_eval_f32': ('eval', [16, 16], 128, 8, 4.0),
    'u1_16x16_nb128_l2hmc_train_f32': ('train', [16, 16], 128, 8, 4.0),
}
DEFAULT_WORKLOAD = 'su3_16x16x16x16_nb64_nlf10_c128'
METRIC = 'link-updates/sec (chains*V*d*Nlf/s)'
SEED = 9992  # conf/config.yaml:11


def peaks() -> tuple[float, str]:
    p = ROOT / 'MEASURED_PEAKS.json'
    if p.exists():
        return float(json.loads(p.read_text())['hbm_gbs']), 'measured'
    return 6650.0, 'fallback'


def links_of(lattice, nb, dim):
    v = 1
    for s in lattice:
        v *= s
    return nb * v * dim


class ClockSampler:
    """nvidia-smi clocks/throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,'
         'clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile('w+', suffix='.csv', delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(
                ['nvidia-smi', f'--id={self.idx}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self) -> dict:
        if self.p is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        time.sleep(0.12)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        for line in self.f.read().strip().splitlines():
            c = [t.strip() for t in line.split(',')]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for n, val in zip(names, c[5:9]):
                if val.lower().startswith('active'):
                    reasons.add(n)
        os.unlink(self.f.name)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'samples': len(sm), 'reasons': sorted(reasons)}


# ---------------------------------------------------------------------------
# reference / cpu_baseline arm
# ---------------------------------------------------------------------------
def reference_sample(workload: str, nsteps: int):
    """bounded sample of the workload for the CPU arm: same lattice, dtype,
    step size and integrator, fewer chains (cost is linear in chains) and, when
    more than 10 trajectories are requested, fewer leapfrog steps per trajectory,
    so that the whole run is <= 100 chain-leapfrog-steps of 16^4 (~1 min on 16
    cores).  Up to 10 trajectories the sample keeps the workload's own N_LF, so
    the two end-point Hamiltonians weigh on the CPU number exactly as they do on
    ours."""
    group, lattice, nb, nlf, dtype, beta = WORKLOADS[workload]
    if group == 'SU3':
        v = 1
        for s in lattice:
            v *= s
        nb_s = max(1, min(nb, 65536 // v))         # 16^4 -> 1 chain, 8^4 -> 16 chains
        nlf_s = max(1, min(nlf, 100 // max(1, nsteps)))   # <= 10 trajectories: the workload's own N_LF
    else:
        nb_s, nlf_s = min(nb, 512), nlf
    return group, lattice, nb_s, nlf_s, dtype, beta


def run_reference_steps(workload: str, steps: int, warmup: int):
    """Times the reference's own `Dynamics.transition_kernel_hmc` (unmodified
    modules under oracle/_ref, loaded by oracle/ref_shim.py) on the host cores.
    Falls back to the numpy oracle port when oracle/_ref did not travel."""
    import numpy as np
    import torch
    group, lattice, nb, nlf, dtype, beta = reference_sample(workload, steps + warmup)
    assert not torch.cuda.is_available(), 'the CPU arm must not see a GPU (the reference moves itself to CUDA)'
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    eps = 1.0 / WORKLOADS[workload][3]
    from oracle import ref_shim
    kind = 'reference' if ref_shim.available() else 'port'
    dim = 4 if group == 'SU3' else 2
    units = links_of(lattice, nb, dim) * nlf
    if kind == 'reference':
        ref = ref_shim.load_reference(torch.float64 if dtype == 'f64' else torch.float32)
        torch.manual_seed(SEED)
        lat = (ref.LatticeSU3 if group == 'SU3' else ref.LatticeU1)(nb, lattice)
        cfg = ref.DynamicsConfig(nchains=nb, group=group, latvolume=lattice, nleapfrog=nlf, eps=eps, eps_hmc=eps,
                                 verbose=False, use_split_xnets=False, use_separate_networks=False,
                                 merge_directions=True)
        dyn = ref.Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
        x = lat.random().detach()
        b = torch.tensor(beta)

        def step():
            v = lat.g.random_momentum(list(cfg.xshape))
            st = ref.State(x=x, v=v, beta=b)
            sp, met = dyn.transition_kernel_hmc(st, eps=eps, nleapfrog=nlf)
            return float(met['acc'].mean())
    else:
        from oracle import su3 as osu3, dynamics as od
        rng = np.random.default_rng(SEED)
        if group == 'SU3':
            full = (nb, 4, *lattice, 3, 3)
            x = osu3.random_su3(rng, full)
            ops_, mom = od.SU3Ops, (lambda: osu3.random_momentum(rng, full))
        else:
            dt = np.float64 if dtype == 'f64' else np.float32
            x = rng.uniform(-np.pi, np.pi, (nb, 2, *lattice)).astype(dt)
            ops_, mom = od.U1Ops, (lambda: rng.standard_normal((nb, 2 * lattice[0] * lattice[1])).astype(dt))

        def step():
            sp, acc = od.transition_kernel_hmc(ops_, od.State(x, mom(), beta), eps, nlf)
            return float(acc.mean())
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt_s = time.perf_counter() - t0
    sample = (f'{group} {"x".join(map(str, lattice))}, {nb} chain(s), {nlf} leapfrog steps per trajectory, '
              f'{dtype}, eps={eps:g}, torch {torch.__version__} CPU, {cores} threads')
    return {'value': units * steps / dt_s, 'unit': 'link-updates/s', 'cores': cores, 'kind': kind,
            'sample': sample}, dt_s / steps * 1e3


def run_reference_l2hmc(workload: str, steps: int, warmup: int):
    """The reference's own SU(3) L2HMC step on the host cores (unmodified modules under
    oracle/_ref): `Dynamics.forward` for the eval workload, forward + LatticeLoss + backward + Adam for
    the training workload; same lattice, N_LF, network config and step size, 2 chains (the cost is
    linear in chains; the 180 M-parameter vnet is built in full), float64 nets: the reference's SU(3)
    path only runs with float64 as torch's default dtype (conf/experiment/su3*.yaml: precision float64)."""
    import numpy as np
    import torch
    mode, lattice, nb, nlf, units, beta = L2HMC_WORKLOADS[workload]
    assert not torch.cuda.is_available(), 'the CPU arm must not see a GPU'
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import ref_shim
    if not ref_shim.available():
        return None, 'oracle/_ref did not travel to this box'
    ref = ref_shim.load_reference(torch.float64)     # the reference's SU(3) path needs float64 as default dtype
    nb_s = 2
    torch.manual_seed(SEED)
    np.random.seed(SEED)
    V = 1
    for s_ in lattice:
        V *= s_
    cfg = ref.DynamicsConfig(nchains=nb_s, group='SU3', latvolume=lattice, nleapfrog=nlf, eps=0.01, eps_hmc=0.01,
                             verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
    ispec = ref.InputSpec(xshape=cfg.xshape, xnet={'x': [4 * V * 8], 'v': [4 * V * 8]},
                          vnet={'x': [4 * V * 8], 'v': [4 * V * 8]})
    ncfg = ref.NetworkConfig(units=[units], activation_fn='tanh', dropout_prob=0.0, use_batch_norm=False)
    nw = ref.NetWeights(x=ref.NetWeight(0., 1., 1.), v=ref.NetWeight(1., 1., 1.))
    lat = ref.LatticeSU3(nb_s, lattice)
    fac = ref.NetworkFactory(input_spec=ispec, network_config=ncfg, conv_config=None, net_weights=nw)
    dyn = ref.Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    x = lat.random().detach()
    b = torch.tensor(beta)
    lcfg = ref.cfgs.LossConfig(use_mixed_loss=True, charge_weight=0.0, rmse_weight=0.1, plaq_weight=0.1)
    loss_fn = ref.LatticeLoss(lat, lcfg)
    opt = torch.optim.Adam([p for p in dyn.parameters() if p.requires_grad], lr=1e-4)

    def step():
        if mode == 'train':
            opt.zero_grad()
        xo, m = dyn((x, b))
        xp = m.pop('mc_states').proposed.x
        loss = loss_fn(x_init=x, x_prop=xp, acc=m['acc'])
        if mode == 'train':
            loss.backward()
            opt.step()
        return float(loss)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt_s = time.perf_counter() - t0
    units_ = nb_s * 4 * V * 2 * nlf
    sample = (f'SU3 {"x".join(map(str, lattice))} L2HMC {mode} step, {nb_s} chains, N_LF {nlf}, units [{units}], '
              f'float64 nets + complex128 lattice, torch {torch.__version__} CPU, {cores} threads')
    return {'value': units_ * steps / dt_s, 'unit': 'link-updates/s', 'cores': cores, 'kind': 'reference',
            'sample': sample}, dt_s / steps * 1e3


def u1_default_configs(mod, nb, lattice, nlf):
    """the default experiment's dynamics / network / conv configs (conf/config.yaml), built from the config classes
    of `mod` (ours or the reference's: field-compatible dataclasses)"""
    cfg = mod.DynamicsConfig(nchains=nb, group='U1', latvolume=lattice, nleapfrog=nlf, eps=0.1, eps_hmc=None,
                             use_ncp=True, verbose=False, eps_fixed=False, use_split_xnets=True, merge_directions=True,
                             use_separate_networks=True)
    ncfg = mod.NetworkConfig(units=[16, 16, 16, 16], activation_fn='leaky_relu', dropout_prob=0.2, use_batch_norm=True)
    ccfg = mod.ConvolutionConfig(filters=[8, 16, 32, 64, 128], sizes=[5, 3, 3, 3, 2], pool=[2, 2, 2, 2, 2])
    return cfg, ncfg, ccfg


def run_reference_u1_l2hmc(workload: str, steps: int, warmup: int):
    """BASELINE cfg 1 on the reference itself (unmodified modules under oracle/_ref), host cores, the FULL config:
    128 chains, `Dynamics.forward` (eval) or forward + LatticeLoss + backward + Adam (train)."""
    import numpy as np
    import torch
    mode, lattice, nb, nlf, beta = U1_L2HMC_WORKLOADS[workload]
    assert not torch.cuda.is_available(), 'the CPU arm must not see a GPU'
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    from oracle import ref_shim
    if not ref_shim.available():
        return None, 'oracle/_ref did not travel to this box'
    ref = ref_shim.load_reference(torch.float32)
    torch.manual_seed(SEED)
    np.random.seed(SEED)
    cfg, ncfg, ccfg = u1_default_configs(ref.cfgs, nb, lattice, nlf)
    V = lattice[0] * lattice[1]
    xdim = 2 * V
    ispec = ref.InputSpec(xshape=cfg.xshape, xnet={'x': [xdim, 2], 'v': [xdim]},     # trainers/trainer.py:292-309
                          vnet={'x': [xdim], 'v': [xdim]})
    lat = ref.LatticeU1(nb, lattice)
    fac = ref.NetworkFactory(input_spec=ispec, network_config=ncfg, conv_config=ccfg, net_weights=ref.NetWeights())
    dyn = ref.Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
    x = lat.random().detach()           # [nb, 2, T, X], as the trainer hands it over (trainer.py:935-938)
    b = torch.tensor(beta)
    loss_fn = ref.LatticeLoss(lat, ref.cfgs.LossConfig(use_mixed_loss=True, charge_weight=0.01))
    opt = torch.optim.Adam([p for p in dyn.parameters() if p.requires_grad], lr=1e-3)
    dyn.train(mode == 'train')

    def step():
        if mode == 'train':
            opt.zero_grad()
            xo, m = dyn((x, b))
            xp = m.pop('mc_states').proposed.x
            loss = loss_fn(x_init=x, x_prop=xp, acc=m['acc'])
            loss.backward()
            opt.step()
        else:      # the reference's eval_step runs with autograd on (its U(1) force IS an autograd call), trainer.py:930-956
            xo, m = dyn((x, b))
            xp = m.pop('mc_states').proposed.x
            loss = loss_fn(x_init=x, x_prop=xp, acc=m['acc'])
        return float(loss)
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(steps):
        step()
    dt_s = time.perf_counter() - t0
    units_ = nb * 2 * V * 2 * nlf
    sample = (f'U1 {"x".join(map(str, lattice))} L2HMC {mode} step, the full {nb} chains, N_LF {nlf}, default experiment '
              f'config (conv stack + units [16,16,16,16]), fp32, torch {torch.__version__} CPU, {cores} threads')
    return {'value': units_ * steps / dt_s, 'unit': 'link-updates/s', 'cores': cores, 'kind': 'reference',
            'sample': sample}, dt_s / steps * 1e3


def main_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    # The reference moves itself to CUDA whenever torch sees a device
    # (l2hmc/__init__.py:45-51, dynamics.py:194-200); this arm times its CPU path.
    os.environ['CUDA_VISIBLE_DEVICES'] = ''
    base, ms = run_reference_steps(args.workload, args.steps, args.warmup)
    group, lattice, nb, nlf, dtype, beta = WORKLOADS[args.workload]
    line = {
        'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'link-updates/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
        'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': 'f64' if dtype == 'f64' else 'f32', 'data': 'synthetic',
        'config': {'workload': args.workload, 'group': group, 'lattice': lattice, 'chains_per_gpu': nb,
                   'nleapfrog': nlf, 'beta': beta, 'sample': base['sample']},
        'cpu_baseline': base,
        'e2e': {'value': base['value'], 'unit': 'link-updates/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))


# ---------------------------------------------------------------------------
# our arm
# ---------------------------------------------------------------------------
class Ctx:
    """process-wide state of the product arm: rank / world / device, the process group, lazily imported package"""

    def __init__(self):
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.rank = int(os.environ.get('RANK', '0'))
        self.world = int(os.environ.get('WORLD_SIZE', '1'))
        self.local = int(os.environ.get('LOCAL_RANK', '0'))
        assert torch.cuda.is_available(), 'bench.py (impl=ours) needs a GPU; there is no CPU fallback'
        torch.cuda.set_device(self.local)
        self.dev = torch.device('cuda', self.local)
        from l2hmc_b200 import dist as l2d
        self.cpus = l2d.bind_rank_to_cores(self.local, self.world)       # before any pinned allocation
        if self.world > 1:
            dist.init_process_group('nccl', device_id=self.dev)
        self.peak, self.peak_kind = peaks()

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def max_over_ranks(self, ms: float) -> float:
        t = self.torch.tensor([ms], dtype=self.torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return float(t)

    def free(self):
        import gc
        gc.collect()
        self.torch.cuda.empty_cache()


def hmc_workload(ctx: Ctx, workload: str, steps: int, warmup: int, thermalise: int = 0, roofline: bool = True,
                 e2e: bool = True, clocks: bool = True, parity: bool = False) -> dict:
    """ONE HMC workload: device-timed trajectories on resident inputs (`value`), the per-kernel roofline, and the
    same metric end to end through `Dynamics.apply_transition_hmc` with host buffers (`e2e`).  Returns the pieces
    of the JSON line; rank 0's copy is the one printed."""
    torch = ctx.torch
    from l2hmc_b200 import _lib, ops
    from l2hmc_b200.configs import DynamicsConfig
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    group, lattice, nb, nlf, dtype, beta = WORKLOADS[workload]
    su3 = group == 'SU3'
    dim = 4 if su3 else 2
    eps = 1.0 / nlf                                           # configs.py:485-487
    units_rank = links_of(lattice, nb, dim) * nlf             # link-updates per trajectory per GPU
    torch.manual_seed(SEED + rank)
    tdt = torch.float64 if dtype == 'f64' else torch.float32
    old_dt = torch.get_default_dtype()
    torch.set_default_dtype(tdt)
    try:
        # synthetic hot-start configuration + momenta, resident in HBM
        cfg = DynamicsConfig(nchains=nb, group=group, latvolume=lattice, nleapfrog=nlf, eps=eps, eps_hmc=eps,
                             verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
        lat = LatticeSU3(nb, lattice) if su3 else LatticeU1(nb, lattice)
        dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
        x = lat.random()
        acc_therm = plaq_therm = None
        if thermalise > 0:
            # hot-start links make the acos / exp arguments atypical; time from a configuration relaxed
            # by N accept/reject HMC trajectories (the reference's `Trainer.warmup`, trainer.py:1699-1744).  The
            # relaxation runs at a fifth of the workload's step size: at eps = 1 / N_LF a 16^4 lattice stops accepting
            # after the first few trajectories and the chain would stay close to the hot start.
            eps_t = eps / 5.0
            accs = []
            with torch.no_grad():
                for _ in range(thermalise):
                    xo_, m_ = dyn.apply_transition_hmc((x, torch.tensor(beta)), eps=eps_t, nleapfrog=nlf)
                    x = xo_.reshape(x.shape)
                    accs.append(m_['acc_mask'].mean())
                acc_therm = float(torch.stack(accs[-20:]).mean())
                plaq_therm = float(lat.plaqs(x).mean())
            x = x.contiguous()
        v = lat.random_momentum()
        field_bytes = x.numel() * x.element_size()

        def traj():
            if su3:
                return ops.su3_hmc_trajectory(x, v, beta, eps, nlf)
            return ops.u1_hmc_trajectory(x, v, beta, eps, nlf, shape=lattice)

        for _ in range(warmup):
            out = traj()
        ctx.barrier()
        sampler = ClockSampler(ctx.local) if (rank == 0 and clocks) else None
        if sampler:
            sampler.start()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        e0.record()
        for _ in range(steps):
            out = traj()
        e1.record()
        ctx.barrier()
        launches = _lib.launch_count() - l0
        clk = sampler.stop() if sampler else None
        ms_max = ctx.max_over_ranks(e0.elapsed_time(e1))
        value = world * units_rank * steps / (ms_max * 1e-3)
        en = out[2]
        assert torch.isfinite(en).all(), 'non-finite energies'
        gb = 864.0 if su3 else 24.0
        res = {
            'workload': workload, 'value': value, 'ms_per_step': ms_max / steps, 'gpu_launches': launches, 'clocks': clk,
            'dtype': 'f64' if dtype == 'f64' else 'f32',
            'config': {'workload': workload, 'group': group, 'lattice': lattice, 'chains_per_gpu': nb,
                       'global_chains': nb * world, 'nleapfrog': nlf, 'eps': eps, 'beta': beta,
                       'start': ('hot (g.random)' if thermalise <= 0 else
                                 f'thermalised ({thermalise} HMC trajectories at eps/5 from a hot start; accept rate over the last 20 = '
                                 f'{acc_therm:.2f}, plaquette {plaq_therm:.4f})'),
                       'parallelism': f'chains sharded over {world} GPU(s), no data-path collective',
                       'l2_policy': f'inputs larger than L2 ({field_bytes / 2**20:.0f} MiB per field per GPU), no flush'
                       if field_bytes > 200 * 2**20 else 'working set fits L2; fields re-read every step (no flush)',
                       'host_cores_bound': ctx.cpus},
            'hbm_model': {'bytes_per_link_update': gb, 'achieved_GBps_per_gpu': value / world * gb / 1e9,
                          'frac_of_peak': value / world * gb / 1e9 / ctx.peak, 'peak_GBps': ctx.peak,
                          'peak_kind': ctx.peak_kind},
        }
        # ---- e2e through the public API with host buffers (before the parity / roofline legs: the reference's own
        # CUDA path leaves the caching allocator fragmented, which cost the e2e loop 13 ms per step when it ran after)
        if e2e:
            res['e2e'] = hmc_e2e(ctx, dyn, x, beta, eps, nlf, units_rank, steps, warmup, su3, tdt)
        # ---- parity of the timed batch: chain 0's proposal against the oracle, after the timed region ----------
        if parity and su3:
            res['parity'] = parity_of_timed_batch(ctx, x, v, out, beta, eps, nlf, lattice)
            ctx.free()
        # ---- per-kernel CUDA-event timing of the dominant kernel ------------------------------------------------
        if roofline:
            if su3:
                res['roofline'] = su3_kernel_roofline(ops, _lib, x, v, lattice, nb, nlf, beta, eps, steps, ctx.peak,
                                                      ctx.peak_kind, ms_max / steps)
            else:
                # whole trajectory is ONE kernel with the state resident in shared memory:
                # algorithmic (streaming-model) bytes vs time; real HBM traffic is 4 field passes
                ach = 24.0 * units_rank / (ms_max / steps * 1e-3) / 1e9
                res['roofline'] = {'bound': 'hbm', 'kernel': 'k_u1_hmc (whole trajectory on-chip)', 'achieved': ach,
                                   'peak': ctx.peak, 'unit': 'GB/s', 'frac': ach / ctx.peak, 'peak_kind': ctx.peak_kind,
                                   'traffic': None,
                                   'note': 'streaming model 24 B/link-update; actual DRAM traffic is 16 B/link per TRAJECTORY'}
        del out, en
        del x, v, dyn, lat
        ctx.free()
        return res
    finally:
        torch.set_default_dtype(old_dt)


def hmc_e2e(ctx: Ctx, dyn, x, beta, eps, nlf, units_rank, steps, warmup, su3, tdt) -> dict:
    """Every step: H2D of that step's links from pinned host memory, the public `Dynamics.apply_transition_hmc`
    call, D2H of the step's results: the accept probabilities AND the new configuration x_out.  Three streams and
    three buffers per direction: the uploads of steps i+1, i+2 and the download of step i-1 run while step i
    computes, so neither copy engine ever waits for a kernel (with two buffers the upload of step i+2 could only
    start when step i had finished: 66.9 ms per step against the 52.9 ms both directions need on this host link,
    `profiles/exp_pcie_bidir.py`) -- as a production sampler feeding independent batches would do."""
    torch = ctx.torch
    dev = ctx.dev
    nb = x.shape[0]
    NBUF = 3
    xh = x.detach().cpu().pin_memory()
    field_bytes = x.numel() * x.element_size()
    xo_h = [torch.empty((nb, x.numel() // nb), dtype=x.dtype).pin_memory() for _ in range(NBUF)]
    acc_h = torch.empty(nb, dtype=torch.float64 if su3 else tdt).pin_memory()
    bt = torch.tensor(beta)
    up, down = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    main = torch.cuda.current_stream(dev)
    xin = [torch.empty_like(x) for _ in range(NBUF)]
    # Full duplex or one direction at a time?  On one GPU both directions together move 2 x 45.7 GB/s against 55.6 / 51.7
    # alone, so duplex wins; with eight ranks behind the VM's shared host path the picture can invert.  One upload and
    # one download of a field are timed both ways on all ranks at once, and the faster policy runs the timed loop.
    probe_out = torch.empty_like(x)

    def probe(duplex: bool) -> float:
        ctx.barrier()
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(main)
        up.wait_stream(main)
        with torch.cuda.stream(up):
            xin[0].copy_(xh, non_blocking=True)
        s_down = down if duplex else up
        s_down.wait_stream(main)
        with torch.cuda.stream(s_down):
            xo_h[0].copy_(probe_out.reshape(xo_h[0].shape), non_blocking=True)
        main.wait_stream(up)
        main.wait_stream(down)
        t1.record(main)
        ctx.barrier()
        return ctx.max_over_ranks(t0.elapsed_time(t1))
    probe(True)
    ms_duplex, ms_serial = probe(True), probe(False)
    duplex = ms_duplex <= ms_serial
    if not duplex:
        down = up                                          # one copy stream: transfers take turns
    del probe_out
    xout = [None] * NBUF
    ready = [torch.cuda.Event() for _ in range(NBUF)]
    freed = [torch.cuda.Event() for _ in range(NBUF)]
    done = [torch.cuda.Event() for _ in range(NBUF)]
    drained = [torch.cuda.Event() for _ in range(NBUF)]

    def stage(i):
        buf = i % NBUF
        with torch.cuda.stream(up):
            up.wait_event(freed[buf])                     # previous user of this input buffer is done
            xin[buf].copy_(xh, non_blocking=True)         # H2D of step i's links
            ready[buf].record(up)

    def run(n):
        for b_ in range(NBUF):
            freed[b_].record(main)
            drained[b_].record(down)
        for i in range(min(NBUF - 1, n)):
            stage(i)
        for i in range(n):
            if i + NBUF - 1 < n:
                stage(i + NBUF - 1)
            buf = i % NBUF
            main.wait_event(ready[buf])
            main.wait_event(drained[buf])                 # step i-3's x_out has left its device buffer
            xo, met = dyn.apply_transition_hmc((xin[buf], bt), eps=eps, nleapfrog=nlf)
            xout[buf] = xo
            freed[buf].record(main)
            done[buf].record(main)
            with torch.cuda.stream(down):
                down.wait_event(done[buf])
                xo_h[buf].copy_(xo, non_blocking=True)    # D2H of step i's new configuration
                acc_h.copy_(met['acc'], non_blocking=True)
                drained[buf].record(down)
        main.wait_stream(down)
        return xo

    with torch.no_grad():
        run(max(1, min(warmup, 2)))
        ctx.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n_e2e = max(steps, 24)                  # pipeline fill (first upload) and drain (last download) amortised
        e2.record()
        run(n_e2e)
        e3.record()
        ctx.barrier()
    ms = ctx.max_over_ranks(e2.elapsed_time(e3))
    # the copy engines alone, same buffers: what the host link gives this rank while all ranks copy at once
    ctx.barrier()
    c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    c0.record()
    for _ in range(2):
        xin[0].copy_(xh, non_blocking=True)
    c1.record()
    ctx.barrier()
    h2d_ms = ctx.max_over_ranks(c0.elapsed_time(c1)) / 2
    return {'value': ctx.world * units_rank * n_e2e / (ms * 1e-3), 'unit': 'link-updates/s',
            'h2d_bytes_per_step': field_bytes, 'd2h_bytes_per_step': field_bytes + acc_h.numel() * acc_h.element_size(),
            'steps': n_e2e, 'ms_per_step': ms / n_e2e,
            'h2d_GBps_per_gpu_all_ranks_copying': field_bytes / (h2d_ms * 1e-3) / 1e9,
            'copy_policy': 'duplex (two copy streams)' if duplex else 'one direction at a time (one copy stream)',
            'copy_probe_ms': {'duplex': ms_duplex, 'serial': ms_serial},
            'api': 'Dynamics.apply_transition_hmc((x_host_pinned -> device, beta)); H2D of steps i+1, i+2 and D2H of '
                   'step i-1 (x_out, acc) overlapped with step i: two copy streams, three buffers per direction'}


def parity_of_timed_batch(ctx: Ctx, x, v, out, beta, eps, nlf, lattice) -> dict:
    """chain 0 of the timed batch, the very tensors the timed trajectories read and wrote, against the oracle on
    same inputs: the reference's own `Dynamics.transition_kernel_hmc` (oracle/_ref) when it travelled, else the
    numpy restatement (oracle/dynamics.py) on the host.  Rank 0 only; runs after the timed region."""
    if ctx.rank != 0:
        return None
    import numpy as np
    torch = ctx.torch
    gx, gv, gen = out[0][:1].cpu().numpy(), out[1][:1].cpu().numpy(), out[2][:1].cpu().numpy()
    t0 = time.perf_counter()
    from oracle import ref_shim
    if ref_shim.available():
        # the reference puts its module constants on CUDA as soon as torch sees a GPU (l2hmc/__init__.py:45-51), so
        # in this process it runs its own ATen/CUDA path (none of our kernels) on the same device
        kind = 'reference (oracle/_ref, its own ATen path on ' + str(x.device) + ')'
        ref = ref_shim.load_reference(torch.float64)
        old = torch.get_default_dtype()
        torch.set_default_dtype(torch.float64)
        try:
            x0, v0 = x[:1].clone(), v[:1].clone()
            lat = ref.LatticeSU3(1, lattice)
            cfg = ref.DynamicsConfig(nchains=1, group='SU3', latvolume=lattice, nleapfrog=nlf, eps=eps, eps_hmc=eps,
                                     verbose=False, use_split_xnets=False, use_separate_networks=False,
                                     merge_directions=True)
            rdyn = ref.Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
            bt = torch.tensor(beta, device=x.device)
            sp, met = rdyn.transition_kernel_hmc(ref.State(x=x0, v=v0, beta=bt), eps=eps, nleapfrog=nlf)
            wx, wv = sp.x.detach().reshape(x0.shape).cpu().numpy(), sp.v.detach().reshape(v0.shape).cpu().numpy()
            wh0 = (lat.action(x0, bt) + lat.g.kinetic_energy(v0)).detach().cpu().numpy()
            wh1 = (lat.action(sp.x.detach().reshape(x0.shape), bt)
                   + lat.g.kinetic_energy(sp.v.detach())).detach().cpu().numpy()
            del sp, met, rdyn, lat
        finally:
            torch.set_default_dtype(old)
    else:
        kind = 'port (oracle/dynamics.py, numpy)'
        from oracle import dynamics as od, su3 as osu3
        xn, vn = x[:1].cpu().numpy(), v[:1].cpu().numpy()
        want, _ = od.transition_kernel_hmc(od.SU3Ops, od.State(xn, vn, beta), eps, nlf)
        wx, wv = want.x, want.v
        wh0 = osu3.action(xn, beta) + osu3.kinetic_energy(vn)
        wh1 = osu3.action(wx, beta) + osu3.kinetic_energy(wv)
    gh0, gh1 = gen[:, 0] + gen[:, 1], gen[:, 2] + gen[:, 3]
    dx, dv = float(np.abs(gx - wx).max()), float(np.abs(gv - wv).max())
    dh = max(float(np.abs(gh0 - wh0).max() / np.abs(wh0).max()), float(np.abs(gh1 - wh1).max() / np.abs(wh1).max()))
    dacc = float(np.abs(np.exp(np.minimum(gh0 - gh1, 0)) - np.exp(np.minimum(wh0 - wh1, 0))).max())
    ok = dx < 1e-12 and dv < 1e-12 and dh < 1e-12 and dacc < 1e-12 * max(1.0, float(np.abs(wh0).max()))
    return {'oracle': kind, 'chains_checked': 1, 'nleapfrog': nlf, 'max_abs_dx': dx, 'max_abs_dv': dv,
            'max_rel_dH': dh, 'max_abs_dacc': dacc, 'tolerance': '1e-12 (links, momenta abs; H rel; acc * max(1,|H|))',
            'ok': bool(ok), 'oracle_seconds': time.perf_counter() - t0}


def gpu_reference_subprocess(workload: str):
    """the reference's OWN `.cuda()` path on this GPU (profiles/time_reference_gpu.py): the like-for-like baseline
    SURVEY 8(d) asks for -- what a user of the reference gets on the same B200 today"""
    try:
        r = subprocess.run([sys.executable, str(ROOT / 'profiles' / 'time_reference_gpu.py'), workload],
                           capture_output=True, text=True, timeout=600)
        d = json.loads([ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1])
        return {'value': d['link_updates_per_s'], 'unit': 'link-updates/s', 'ms_per_step': d['ms_per_trajectory'],
                'sample': f"{d['chains']} chains of the same lattice / N_LF / dtype / eps on {d['device']} (the reference "
                          f"materialises ~40 field-sized temporaries per force evaluation; peak {d['peak_mem_GB']:.1f} GB), "
                          'per-trajectory wall time with synchronize', 'impl': d['impl']}
    except Exception as e:
        return {'value': None, 'unit': 'link-updates/s', 'sample': f'failed: {type(e).__name__}: {e}'}


def main_ours(args):
    ctx = Ctx()
    torch = ctx.torch
    head = hmc_workload(ctx, args.workload, args.steps, args.warmup, thermalise=args.thermalise,
                        parity=not args.no_parity)
    group, lattice, nb, nlf, dtype, beta = WORKLOADS[args.workload]
    full = (args.workload == DEFAULT_WORKLOAD) and not args.headline_only
    line = {
        'metric': METRIC, 'value': head['value'], 'unit': 'link-updates/s', 'n_gpus': ctx.world, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': head['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': head['dtype'], 'data': 'synthetic', 'config': head['config'],
        'hbm_model': head['hbm_model'], 'roofline': head.get('roofline'), 'cpu_baseline': None,
        'e2e': head.get('e2e'), 'gpu_launches': head['gpu_launches'], 'clocks': head['clocks'],
        'parity': head.get('parity'),
    }
    if full:
        # ---- thermalised start (trainer.py:1699-1744): same workload from a relaxed configuration ----
        th = hmc_workload(ctx, args.workload, max(2, min(args.steps, 5)), 3, thermalise=args.thermalise_n, roofline=False,
                          e2e=False, clocks=False)
        line['thermalised'] = {'value': th['value'], 'ms_per_step': th['ms_per_step'], 'start': th['config']['start']}
        # ---- the other BASELINE configs, same run, same process group ------------------------------------------------
        sec = {}
        for w in ('su3_8x8x8x8_nb256_nlf10_c128', 'u1_64x64_nb4096_nlf10_f32'):
            r = hmc_workload(ctx, w, max(3, min(args.steps, 10)), 3, clocks=False)
            sec[w] = {'value': r['value'], 'ms_per_step': r['ms_per_step'], 'unit': 'link-updates/s',
                      'roofline_frac': r['roofline']['frac'], 'roofline_kernel': r['roofline']['kernel'],
                      'hbm_model_frac': r['hbm_model']['frac_of_peak'], 'e2e': r['e2e']['value'],
                      'gpu_launches': r['gpu_launches'], 'chains_per_gpu': r['config']['chains_per_gpu']}
        for w in ('su3_8x8x8x8_nb256_l2hmc_eval_bf16', 'su3_8x8x8x8_nb32_l2hmc_train_bf16'):
            # both steps replay from CUDA graphs at every N (the eager steps are host-bound once eight ranks share the
            # node's cores: 48 ms against 25 ms); multi-rank training = graph (forward + backward + pack), one eager
            # NCCL all-reduce of the flat bf16 gradient bucket, graph (unpack + clip + Adam)
            r = l2hmc_workload(ctx, w, max(3, min(args.steps, 5)), 3, cuda_graphs=True, clocks=False)
            sec[w] = {'value': r['value'], 'ms_per_step': r['ms_per_step'], 'unit': 'link-updates/s',
                      'roofline_frac': r['roofline']['frac'], 'roofline_kernel': r['roofline']['kernel'],
                      'e2e': r['e2e']['value'], 'gpu_launches': r['gpu_launches'],
                      'chains_per_gpu': r['config']['chains_per_gpu'], 'parallelism': r['config']['parallelism'],
                      'cuda_graphs': r['config']['cuda_graphs'], 'ms_each_step': r['ms_each_step'],
                      'grad_allreduce': r.get('grad_allreduce')}
        if ctx.world == 1:
            # BASELINE cfg 1 (the reference's default experiment; the CPU arm runs it at full size: `--impl reference
            # --workload u1_16x16_nb128_l2hmc_train_f32`)
            for w in sorted(U1_L2HMC_WORKLOADS):
                r = u1_l2hmc_workload(ctx, w, max(3, min(args.steps, 10)), 3, cuda_graphs=True, clocks=False)
                sec[w] = {'value': r['value'], 'ms_per_step': r['ms_per_step'], 'unit': 'link-updates/s',
                          'roofline_frac': None, 'roofline_kernel': None, 'e2e': r['e2e']['value'],
                          'gpu_launches': r['gpu_launches'], 'chains_per_gpu': r['config']['chains_per_gpu'],
                          'cuda_graphs': True, 'ms_each_step': r['ms_each_step'], 'note': r['config']['l2_policy']}
        line['secondary'] = sec
    if ctx.rank == 0:
        if ctx.world == 1 and not args.no_cpu_baseline:
            line['cpu_baseline'] = cpu_baseline_subprocess(args.workload)
            if full:
                line['gpu_reference'] = gpu_reference_subprocess(args.workload)
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def l2hmc_workload(ctx: Ctx, workload: str, steps: int, warmup: int, cuda_graphs: bool = False,
                   clocks: bool = True) -> dict:
    """SU(3) L2HMC eval / training step through the public Trainer API (BASELINE cfg 3 secondary / cfg 5)."""
    import numpy as np
    torch = ctx.torch
    rank, world, dev = ctx.rank, ctx.world, ctx.dev
    from l2hmc_b200 import _lib, ops
    from l2hmc_b200.configs import (DynamicsConfig, LossConfig, NetWeight, NetWeights, NetworkConfig,
                                    get_input_spec)
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.su3.pytorch.lattice import LatticeSU3
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    from l2hmc_b200.trainers.pytorch.trainer import Trainer
    mode, lattice, nb, nlf, units, beta = L2HMC_WORKLOADS[workload]
    old_dt = torch.get_default_dtype()
    torch.manual_seed(SEED)            # identical initial weights on every rank (Trainer also broadcasts rank 0's)
    np.random.seed(SEED)
    torch.set_default_dtype(torch.float32)
    try:
        cfg = DynamicsConfig(nchains=nb, group='SU3', latvolume=lattice, nleapfrog=nlf, eps=0.01, eps_hmc=0.01,
                             verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
        fac = NetworkFactory(input_spec=get_input_spec(cfg),
                             network_config=NetworkConfig(units=[units], activation_fn='tanh', dropout_prob=0.0,
                                                          use_batch_norm=False),
                             conv_config=None, net_weights=NetWeights(x=NetWeight(0., 1., 1.), v=NetWeight(1., 1., 1.)),
                             build_unused_su3_xnet=False)
        lat = LatticeSU3(nb, lattice)
        dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
        tr = Trainer(dyn, LossConfig(use_mixed_loss=True, charge_weight=0.0, rmse_weight=0.1, plaq_weight=0.1), lr=1e-4,
                     clip_val=1.0, autocast_dtype=torch.bfloat16, grad_bucket_dtype=torch.bfloat16,
                     cuda_graphs=cuda_graphs)
        torch.manual_seed(SEED + 1 + rank)   # per-rank chains
        x = lat.random().to(torch.complex128)
        bt = torch.tensor(beta)
        V = 1
        for s_ in lattice:
            V *= s_
        units_rank = nb * 4 * V * 2 * nlf     # link-updates per step per GPU (nlf forward + nlf backward layers)

        def step(xin):
            if mode == 'eval':
                with torch.no_grad(), torch.autocast('cuda', dtype=torch.bfloat16):
                    return tr.eval_step((xin, bt))
            return tr.train_step((xin, bt))

        for _ in range(warmup):
            step(x)
        ctx.barrier()
        sampler = ClockSampler(ctx.local) if (rank == 0 and clocks) else None
        if sampler:
            sampler.start()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        e0.record()
        marks = [e0]
        for _ in range(steps):
            xo, met = step(x)
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        e1.record()
        ctx.barrier()
        launches = _lib.launch_count() - l0
        if cuda_graphs:       # a replay does not pass through the host-side counter: kernels recorded at capture
            launches = steps * int(tr.graph_launches.get(mode, 0))
        clk = sampler.stop() if sampler else None
        ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
        per_step = [a_.elapsed_time(b_) for a_, b_ in zip(marks[:-1], marks[1:])]     # this rank's steps, one by one
        value = world * units_rank / (ms * 1e-3)
        assert torch.isfinite(met['loss']), 'non-finite loss'
        # e2e: links from pinned host memory every step, loss read back to the host
        xh = x.cpu().pin_memory()
        loss_h = torch.empty((), dtype=met['loss'].dtype).pin_memory()
        xin = torch.empty_like(x)
        n_e2e = max(2, min(steps, 5))
        ctx.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(n_e2e):
            xin.copy_(xh, non_blocking=True)
            xo, met = step(xin)
            loss_h.copy_(met['loss'], non_blocking=True)
        e3.record()
        ctx.barrier()
        e2e_val = world * units_rank * n_e2e / (ctx.max_over_ranks(e2.elapsed_time(e3)) * 1e-3)
        # roofline of the tensor-core kernel of this path: k_heads_vupdate, timed alone with CUDA events
        vnet = dyn._get_vnet(0)
        pack = vnet.heads_pack()
        xdim = pack.xdim
        z = torch.tanh(torch.randn(nb, units, device=dev)).to(torch.bfloat16)
        vv = lat.random_momentum().reshape(nb, xdim)
        ff = lat.random_momentum().reshape(nb, xdim)
        for _ in range(3):
            ops.su3_heads_vupdate(z, pack, vv, ff, 0.01, 1)
        evs = []
        for _ in range(10):
            a_, b_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a_.record()
            ops.su3_heads_vupdate(z, pack, vv, ff, 0.01, 1)
            b_.record()
            evs.append((a_, b_))
        torch.cuda.synchronize()
        ms_k = statistics.mean(a_.elapsed_time(b_) for a_, b_ in evs)
        algo = 3.0 * 16 * nb * xdim + 3.0 * 2 * xdim * units          # v r, F r, v' w (complex128) + bf16 weights once
        ach = algo / (ms_k * 1e-3) / 1e9
        flops = 2.0 * 3 * nb * xdim * units
        res = {
            'workload': workload, 'value': value, 'ms_per_step': ms, 'gpu_launches': launches, 'clocks': clk,
            'ms_each_step': [round(t_, 3) for t_ in per_step],
            'dtype': 'f64 lattice + bf16 nets (fp32 accumulate)',
            'config': {'workload': workload, 'group': 'SU3', 'lattice': lattice, 'chains_per_gpu': nb,
                       'global_chains': nb * world, 'nleapfrog': nlf, 'units': [units], 'beta': beta, 'step': mode,
                       'cuda_graphs': bool(cuda_graphs),
                       'start': 'hot (g.random), random-init weights',
                       'parallelism': (f'chains sharded over {world} GPU(s); '
                                       + (('NN gradients averaged over the ranks in a flat bf16 bucket: '
                                           + ('one NCCL all-reduce between the two CUDA graphs of the step' if cuda_graphs
                                              else 'per-matrix NCCL all-reduces started from inside backward'))
                                          if mode == 'train' else 'no collective')),
                       'l2_policy': 'fields + weights (~1 GB) exceed L2; no flush'},
            'roofline': {'bound': 'hbm', 'kernel': 'k_heads_vupdate (tcgen05 heads GEMM + momentum update)',
                         'achieved': ach, 'peak': ctx.peak, 'unit': 'GB/s', 'frac': ach / ctx.peak,
                         'peak_kind': ctx.peak_kind, 'traffic': None, 'algorithmic_bytes_per_launch': algo,
                         'avg_launch_ms': ms_k, 'tensor_tflops': flops / (ms_k * 1e-3) / 1e12,
                         'note': 'algorithmic bytes = v, F read + v\' written (complex128) + the bf16 head weights once; '
                                 'the GEMM is 0.2 % of the bf16 tensor peak by construction (HBM-bound op)'},
            'e2e': {'value': e2e_val, 'unit': 'link-updates/s', 'h2d_bytes_per_step': x.numel() * x.element_size(),
                    'd2h_bytes_per_step': loss_h.element_size(), 'steps': n_e2e,
                    'api': f'Trainer.{mode}_step((x_host_pinned -> device, beta))'},
            'grad_allreduce': getattr(tr, 'allreduce_info', lambda: None)() if mode == 'train' else None,
        }
        del tr, dyn, fac, lat, x, xin, vv, ff, pack, vnet, xo, met
        ctx.free()
        return res
    finally:
        torch.set_default_dtype(old_dt)


def u1_l2hmc_workload(ctx: Ctx, workload: str, steps: int, warmup: int, cuda_graphs: bool = True,
                      clocks: bool = True) -> dict:
    """BASELINE cfg 1 -- the reference's default experiment (U(1) 16x16, 128 chains, N_LF 8, conv + dense nets, fp32)
    -- through the public Trainer API: eval_step / train_step, every layer on the hand-written kernels (conv stack as
    periodic gather + tcgen05 GEMM, fp32 Linears through bf16x3 splits).  At this size the step is launch-latency
    bound (a few thousand small kernels), so it is timed as the CUDA-graph replay `Trainer(cuda_graphs=True)` gives."""
    import numpy as np
    torch = ctx.torch
    rank, world = ctx.rank, ctx.world
    from l2hmc_b200 import _lib, configs
    from l2hmc_b200.dynamics.pytorch.dynamics import Dynamics
    from l2hmc_b200.lattice.u1.pytorch.lattice import LatticeU1
    from l2hmc_b200.network.pytorch.network import NetworkFactory
    from l2hmc_b200.trainers.pytorch.trainer import Trainer
    mode, lattice, nb, nlf, beta = U1_L2HMC_WORKLOADS[workload]
    old_dt = torch.get_default_dtype()
    torch.manual_seed(SEED)
    np.random.seed(SEED)
    torch.set_default_dtype(torch.float32)
    try:
        cfg, ncfg, ccfg = u1_default_configs(configs, nb, lattice, nlf)
        fac = NetworkFactory(input_spec=configs.get_input_spec(cfg), network_config=ncfg, conv_config=ccfg,
                             net_weights=None)
        lat = LatticeU1(nb, lattice)
        dyn = Dynamics(potential_fn=lat.action, config=cfg, network_factory=fac)
        tr = Trainer(dyn, configs.LossConfig(use_mixed_loss=True, charge_weight=0.01), lr=1e-3, clip_val=1.0,
                     cuda_graphs=cuda_graphs)
        torch.manual_seed(SEED + 1 + rank)
        x = lat.random()
        bt = torch.tensor(beta)
        units_rank = nb * 2 * lattice[0] * lattice[1] * 2 * nlf

        def step(xin):
            return tr.eval_step((xin, bt)) if mode == 'eval' else tr.train_step((xin, bt))

        for _ in range(warmup):
            step(x)
        ctx.barrier()
        sampler = ClockSampler(ctx.local) if (rank == 0 and clocks) else None
        if sampler:
            sampler.start()
        l0 = _lib.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ctx.barrier()
        e0.record()
        marks = [e0]
        for _ in range(steps):
            xo, met = step(x)
            marks.append(torch.cuda.Event(enable_timing=True))
            marks[-1].record()
        e1.record()
        ctx.barrier()
        launches = _lib.launch_count() - l0
        if cuda_graphs:
            launches = steps * int(tr.graph_launches.get(mode, 0))
        clk = sampler.stop() if sampler else None
        ms = ctx.max_over_ranks(e0.elapsed_time(e1)) / steps
        per_step = [a_.elapsed_time(b_) for a_, b_ in zip(marks[:-1], marks[1:])]
        assert torch.isfinite(torch.as_tensor(met['loss'])), 'non-finite loss'
        # e2e: links from pinned host memory every step, loss read back to the host
        xh = x.cpu().pin_memory()
        loss_h = torch.empty((), dtype=torch.float32).pin_memory()
        xin = torch.empty_like(x)
        ctx.barrier()
        e2, e3 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e2.record()
        for _ in range(steps):
            xin.copy_(xh, non_blocking=True)
            xo, met = step(xin)
            loss_h.copy_(torch.as_tensor(met['loss']).to(torch.float32), non_blocking=True)
        e3.record()
        ctx.barrier()
        e2e_val = world * units_rank * steps / (ctx.max_over_ranks(e2.elapsed_time(e3)) * 1e-3)
        res = {
            'workload': workload, 'value': world * units_rank / (ms * 1e-3), 'ms_per_step': ms,
            'ms_each_step': [round(t_, 3) for t_ in per_step],
            'gpu_launches': launches, 'clocks': clk, 'dtype': 'f32',
            'config': {'workload': workload, 'group': 'U1', 'lattice': lattice, 'chains_per_gpu': nb,
                       'global_chains': nb * world, 'nleapfrog': nlf, 'beta': beta, 'step': mode,
                       'network': 'conf/config.yaml defaults: conv [8,16,32,64,128] / sizes [5,3,3,3,2] / pool 2, units '
                                  '[16,16,16,16], leaky_relu, dropout 0.2, batch norm, separate + split networks',
                       'cuda_graphs': bool(cuda_graphs), 'start': 'hot (g.random), random-init weights',
                       'parallelism': f'chains sharded over {world} GPU(s)',
                       'l2_policy': 'working set (< 100 MB) fits L2: the step is launch-latency bound, not HBM bound'},
            'roofline': None,
            'e2e': {'value': e2e_val, 'unit': 'link-updates/s', 'h2d_bytes_per_step': x.numel() * x.element_size(),
                    'd2h_bytes_per_step': 4, 'steps': steps, 'api': f'Trainer.{mode}_step((x_host_pinned -> device, beta))'},
        }
        del tr, dyn, fac, lat, x, xin, xo, met
        ctx.free()
        return res
    finally:
        torch.set_default_dtype(old_dt)


def main_l2hmc(args):
    ctx = Ctx()
    if args.workload in U1_L2HMC_WORKLOADS:
        r = u1_l2hmc_workload(ctx, args.workload, args.steps, args.warmup, cuda_graphs=True)
        r['grad_allreduce'] = None
    else:
        r = l2hmc_workload(ctx, args.workload, args.steps, args.warmup, cuda_graphs=args.cuda_graphs)
    if ctx.rank == 0:
        line = {
            'metric': METRIC, 'value': r['value'], 'unit': 'link-updates/s', 'n_gpus': ctx.world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': r['ms_per_step'], 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': r['dtype'], 'data': 'synthetic', 'config': r['config'],
            'roofline': r['roofline'],
            'cpu_baseline': (cpu_baseline_subprocess(args.workload) if (ctx.world == 1 and not args.no_cpu_baseline)
                             else None),
            'e2e': r['e2e'], 'gpu_launches': r['gpu_launches'], 'clocks': r['clocks'],
            'grad_allreduce': r['grad_allreduce'],
        }
        print(json.dumps(line))
    if ctx.world > 1:
        ctx.dist.destroy_process_group()


def cpu_baseline_subprocess(workload: str):
    """the reference's CPU path, in a child process that cannot see the GPU"""
    env = dict(os.environ, CUDA_VISIBLE_DEVICES='', RANK='0', WORLD_SIZE='1')
    for k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS'):
        env.pop(k, None)
    try:
        r = subprocess.run([sys.executable, str(ROOT / 'bench.py'), '--impl', 'reference', '--workload', workload,
                            '--steps', '1', '--warmup', '1'], env=env, capture_output=True, text=True, timeout=900)
        line = [ln for ln in r.stdout.splitlines() if ln.startswith('{')][-1]
        return json.loads(line)['cpu_baseline']
    except Exception as e:  # the baseline is a reported number, never a reason to lose the GPU result
        return {'value': None, 'unit': 'link-updates/s', 'cores': os.cpu_count(), 'kind': 'reference',
                'sample': f'failed: {type(e).__name__}: {e}'}


def su3_kernel_roofline(ops, _lib, x, v, lattice, nb, nlf, beta, eps, steps, peak, peak_kind, ms_traj):
    """Re-runs the same trajectories kernel by kernel (the C ABI's planar step
    entry points: exactly the launches l2b_su3_hmc_trajectory issues) with CUDA
    events between launches on the launching stream.  Dominant kernel: the fused
    leapfrog step k_force<..., DRIFT=true> (staples + TAH + kick + exp + link update)."""
    import torch
    dims = _lib.dims4(lattice)
    F64 = _lib.L2B_F64
    nws = _lib.su3_ws_bytes(nb, lattice)
    ws = torch.empty(nws, dtype=torch.uint8, device=x.device)
    Ua, Ub, P = torch.empty_like(x), torch.empty_like(x), torch.empty_like(x)
    st = ops._stream()
    p = ops._ptr
    _lib.call('l2b_su3_aos_to_soa', p(x), p(Ua), nb, dims, F64, st)
    _lib.call('l2b_su3_aos_to_soa', p(v), p(P), nb, dims, F64, st)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    t_step, t_last = [], []
    for _ in range(max(1, min(steps, 3))):
        for k in range(nlf):
            a, b = ev(), ev()
            a.record()
            _lib.call('l2b_su3_force_kick_drift_planar', p(Ua), p(P), p(Ub), float(beta),
                      (0.5 if k == 0 else 1.0) * eps, float(eps), None, nb, dims, F64, p(ws), nws, st)
            b.record()
            t_step.append((a, b))
            Ua, Ub = Ub, Ua
        a, b = ev(), ev()
        a.record()
        _lib.call('l2b_su3_force_kick_planar', p(Ua), p(P), float(beta), 0.5 * eps, None, nb, dims, F64, p(ws), nws, st)
        b.record()
        t_last.append((a, b))
    torch.cuda.synchronize()
    ms_s = statistics.mean(a.elapsed_time(b) for a, b in t_step)
    ms_l = statistics.mean(a.elapsed_time(b) for a, b in t_last)
    links = x.numel() // 9
    bytes_model = 864.0 * links      # 6 link-sized transfers per link-update
    bytes_moved = 576.0 * links      # what the fused step has to move: r U, r P, w P, w U'
    ach = bytes_model / (ms_s * 1e-3) / 1e9
    return {'bound': 'hbm', 'kernel': 'k_force_ep<32,3,DRIFT,3> (one fused leapfrog step: staples + TAH + kick + exp + link update)',
            'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'peak_kind': peak_kind,
            # dram__bytes_read.sum + dram__bytes_write.sum of ONE launch at 16^4 x 64 (ncu --set full,
            # profiles/r2_force_ep_ncu_full.md: 10.04 GB vs 9.66 GB the kernel has to move)
            'traffic': 10.043e9 if (list(lattice) == [16, 16, 16, 16] and nb == 64) else None,
            'algorithmic_bytes_per_launch': bytes_model, 'avg_launch_ms': ms_s,
            'share_of_step': nlf * ms_s / ms_traj,
            'note': ('algorithmic bytes = SURVEY 8(d) streaming model, 864 B per link-update (6 transfers); '
                     'the fused kernel itself moves 4 transfers = 576 B/link'),
            'achieved_moved_bytes': bytes_moved / (ms_s * 1e-3) / 1e9,
            'frac_moved_bytes': bytes_moved / (ms_s * 1e-3) / 1e9 / peak,
            'other_kernels': {'k_force (final half kick, no drift)': {
                'avg_launch_ms': ms_l, 'achieved': 432.0 * links / (ms_l * 1e-3) / 1e9,
                'frac': 432.0 * links / (ms_l * 1e-3) / 1e9 / peak}}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', choices=['ours', 'reference'], default='ours')
    ap.add_argument('--workload', choices=sorted(WORKLOADS) + sorted(L2HMC_WORKLOADS) + sorted(U1_L2HMC_WORKLOADS),
                    default=DEFAULT_WORKLOAD)
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--thermalise', type=int, default=0,
                    help='HMC trajectories run before timing (default 0: hot start, as the reference\'s g.random)')
    ap.add_argument('--thermalise-n', type=int, default=100,
                    help='trajectories of the `thermalised` entry the default run adds next to the hot-start headline')
    ap.add_argument('--headline-only', action='store_true',
                    help='skip the thermalised / secondary (BASELINE cfg 2, 3, 5) / gpu_reference entries')
    ap.add_argument('--no-parity', action='store_true',
                    help='skip the oracle check of one chain of the timed batch')
    ap.add_argument('--cuda-graphs', action='store_true',
                    help='L2HMC workloads: run the Trainer step functions as CUDA graphs (Trainer(cuda_graphs=True))')
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == 'ours':
        args.warmup = 3
    if args.impl == 'reference':
        if args.workload in U1_L2HMC_WORKLOADS:
            if int(os.environ.get('RANK', '0')) != 0:
                return
            os.environ['CUDA_VISIBLE_DEVICES'] = ''
            base, ms = run_reference_u1_l2hmc(args.workload, max(1, min(args.steps, 3)), min(args.warmup, 1))
            if base is None:
                print(json.dumps({'impl': 'reference', 'unavailable': ms}))
                return
            mode, lattice, nb, nlf, beta = U1_L2HMC_WORKLOADS[args.workload]
            print(json.dumps({
                'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'link-updates/s',
                'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                'config': {'workload': args.workload, 'group': 'U1', 'lattice': lattice, 'chains_per_gpu': nb,
                           'nleapfrog': nlf, 'beta': beta, 'step': mode, 'sample': base['sample']},
                'cpu_baseline': base,
                'e2e': {'value': base['value'], 'unit': 'link-updates/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))
            return
        if args.workload in L2HMC_WORKLOADS:
            if int(os.environ.get('RANK', '0')) != 0:
                return
            os.environ['CUDA_VISIBLE_DEVICES'] = ''
            base, ms = run_reference_l2hmc(args.workload, max(1, min(args.steps, 2)), min(args.warmup, 1))
            if base is None:
                print(json.dumps({'impl': 'reference', 'unavailable': ms}))
                return
            mode, lattice, nb, nlf, units, beta = L2HMC_WORKLOADS[args.workload]
            print(json.dumps({
                'impl': 'reference', 'metric': METRIC, 'value': base['value'], 'unit': 'link-updates/s',
                'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': ms,
                'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f64 lattice + f64 nets',
                'data': 'synthetic',
                'config': {'workload': args.workload, 'group': 'SU3', 'lattice': lattice, 'chains_per_gpu': nb,
                           'nleapfrog': nlf, 'units': [units], 'beta': beta, 'step': mode, 'sample': base['sample']},
                'cpu_baseline': base,
                'e2e': {'value': base['value'], 'unit': 'link-updates/s', 'h2d_bytes_per_step': 0,
                        'd2h_bytes_per_step': 0}, 'gpu_launches': 0}))
            return
        main_reference(args)
    elif args.workload in L2HMC_WORKLOADS or args.workload in U1_L2HMC_WORKLOADS:
        main_l2hmc(args)
    else:
        main_ours(args)


if __name__ == '__main__':
    main()

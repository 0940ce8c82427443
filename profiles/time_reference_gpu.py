"""The reference's OWN PyTorch path on the GPU (what a user of saforem2/l2hmc-qcd has today on this hardware):
`Dynamics.transition_kernel_hmc` of the unmodified modules under oracle/_ref, which move themselves to CUDA when
torch sees a device (l2hmc/__init__.py:45-51).  SURVEY 8(d) asks for this like-for-like GPU baseline next to the
CPU one that `bench.py --impl reference` reports.  Bounded sample: the reference materialises ~40 field-sized
temporaries per force evaluation, so chains are capped (16^4: 4 chains, 8^4: 32 chains).
Usage: python profiles/time_reference_gpu.py [workload ...]        (workload names as in bench.py)"""
import json
import sys
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402
from oracle import ref_shim  # noqa: E402

assert torch.cuda.is_available(), 'needs a GPU'
assert ref_shim.available(), 'oracle/_ref did not travel'
names = sys.argv[1:] or ['su3_16x16x16x16_nb64_nlf10_c128', 'su3_8x8x8x8_nb256_nlf10_c128', 'u1_64x64_nb4096_nlf10_f32']
if len(names) > 1:
    # the reference freezes dtype-typed constants at import (group/su3/pytorch/utils.py:28-36): one process per workload
    import subprocess
    for name in names:
        subprocess.run([sys.executable, __file__, name], check=False)
    sys.exit(0)
for name in names:
    group, lattice, nb, nlf, dtype, beta = bench.WORKLOADS[name]
    ref = ref_shim.load_reference(torch.float64 if dtype == 'f64' else torch.float32)
    V = 1
    for s in lattice:
        V *= s
    nb_s = nb if group == 'U1' else max(1, min(nb, (1 << 18) // V))          # 16^4 -> 4 chains, 8^4 -> 64 -> cap 32
    nb_s = min(nb_s, 32) if group == 'SU3' else nb_s
    eps = 1.0 / nlf
    torch.manual_seed(bench.SEED)
    lat = (ref.LatticeSU3 if group == 'SU3' else ref.LatticeU1)(nb_s, lattice)
    cfg = ref.DynamicsConfig(nchains=nb_s, group=group, latvolume=lattice, nleapfrog=nlf, eps=eps, eps_hmc=eps,
                             verbose=False, use_split_xnets=False, use_separate_networks=False, merge_directions=True)
    dyn = ref.Dynamics(potential_fn=lat.action, config=cfg, network_factory=None)
    x = lat.random().detach()
    b = torch.tensor(beta, device=x.device)

    def step():
        v = lat.g.random_momentum(list(cfg.xshape))
        sp, met = dyn.transition_kernel_hmc(ref.State(x=x, v=v, beta=b), eps=eps, nleapfrog=nlf)
        return float(met['acc'].mean())
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    n = 3
    t0 = time.perf_counter()
    for _ in range(n):
        step()
    torch.cuda.synchronize()
    dt = (time.perf_counter() - t0) / n
    units = nb_s * V * (4 if group == 'SU3' else 2) * nlf
    print(json.dumps({'workload': name, 'impl': 'reference on the GPU (its own .cuda() path)', 'device': str(x.device),
                      'chains': nb_s, 'ms_per_trajectory': dt * 1e3, 'link_updates_per_s': units / dt,
                      'peak_mem_GB': torch.cuda.max_memory_allocated() / 2**30}))

"""What the host link gives the e2e leg of bench.py: H2D alone, D2H alone, both at once (two copy streams), for one
field of the headline workload (2.42 GB, pinned).  Prints one JSON line."""
import json
import sys

import torch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2415919104
dev = torch.device('cuda', 0)
h_in = torch.empty(n, dtype=torch.uint8).pin_memory()
h_out = torch.empty(n, dtype=torch.uint8).pin_memory()
d_in = torch.empty(n, dtype=torch.uint8, device=dev)
d_out = torch.empty(n, dtype=torch.uint8, device=dev)
up, down = torch.cuda.Stream(), torch.cuda.Stream()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    up.synchronize()
    down.synchronize()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def h2d():
    d_in.copy_(h_in, non_blocking=True)


def d2h():
    h_out.copy_(d_out, non_blocking=True)


def both():
    main = torch.cuda.current_stream()
    up.wait_stream(main)
    down.wait_stream(main)
    with torch.cuda.stream(up):
        d_in.copy_(h_in, non_blocking=True)
    with torch.cuda.stream(down):
        h_out.copy_(d_out, non_blocking=True)
    main.wait_stream(up)
    main.wait_stream(down)


res = {'bytes': n}
for name, fn in (('h2d', h2d), ('d2h', d2h), ('both', both)):
    ms = timed(fn)
    res[name + '_ms'] = ms
    res[name + '_GBps_per_direction'] = n / ms / 1e6
print(json.dumps(res))

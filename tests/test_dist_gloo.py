"""CPU tier: world_size-2 `gloo` test of the N>1 host logic (chain sharding,
observable gather, flat gradient all-reduce).  No GPU kernels are involved."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from l2hmc_b200 import dist as l2d


def test_shard_bounds_cover_exactly():
    for n in (1, 7, 64, 512, 513):
        for world in (1, 2, 3, 8):
            spans = [l2d.shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        l2d.shard_bounds(8, 2, 2)
    assert l2d.rank_seed(9992, 1, 1) == 9992 * 4


def _free_port():
    with socket.socket() as s:
        s.bind(('127.0.0.1', 0))
        return s.getsockname()[1]


def _worker(rank, world, port, nchains, q):
    os.environ.update(MASTER_ADDR='127.0.0.1', MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = l2d.init('gloo')
    assert (r, w) == (rank, world)
    full = torch.arange(nchains * 3, dtype=torch.float64).reshape(nchains, 3)
    mine = l2d.shard_chains(full, rank, world)
    # "observable" computed shard-locally, gathered back in global chain order
    got = l2d.gather_chains(mine.sum(1), nchains)
    ok_gather = torch.equal(got, full.sum(1))
    # gradient all-reduce: mean over ranks, unused parameter skipped
    a = torch.nn.Parameter(torch.zeros(5))
    b = torch.nn.Parameter(torch.zeros(2, 2))
    unused = torch.nn.Parameter(torch.zeros(3))
    a.grad = torch.full((5,), float(rank + 1))
    b.grad = torch.full((2, 2), float(10 * (rank + 1)))
    n = l2d.allreduce_mean_grads([a, unused, b])
    mean = sum(range(1, world + 1)) / world
    ok_grad = (n == 9 and torch.allclose(a.grad, torch.full((5,), mean))
               and torch.allclose(b.grad, torch.full((2, 2), 10 * mean)) and unused.grad is None)
    # GradBucket: fixed parameter list (a parameter without a gradient on this rank counts as zero), one slice
    # reduced early (what the deferred head GEMMs do), the rest in finish(); every rank ends with the same mean
    torch.manual_seed(3)
    ps = [torch.nn.Parameter(torch.zeros(s_)) for s_ in ((4, 3), (7,), (), (2, 5))]
    ok_bucket = True
    for bdt in (None, torch.bfloat16):
        bucket = l2d.GradBucket(ps, bdt)
        bucket.begin()
        for pp in ps:
            pp.grad = None
        big = ps[0]
        bucket.view(big).copy_(torch.full((4, 3), float(rank + 1)))      # produced straight into the bucket
        bucket.reduce_async(big)
        ps[1].grad = torch.arange(7, dtype=torch.float32) * (rank + 1)
        ps[2].grad = torch.tensor(2.0 * (rank + 1))
        if rank == 0:
            ps[3].grad = torch.ones(2, 5)                               # only rank 0 has this gradient
        ncalls = bucket.finish()
        ok_bucket &= ncalls == 2 and bucket.last['early_calls'] == 1
        ok_bucket &= torch.allclose(ps[0].grad, torch.full((4, 3), mean))
        ok_bucket &= torch.allclose(ps[1].grad, torch.arange(7, dtype=torch.float32) * mean)
        ok_bucket &= torch.allclose(ps[2].grad, torch.tensor(2.0 * mean))
        ok_bucket &= torch.allclose(ps[3].grad, torch.full((2, 5), 1.0 / world))
        ok_bucket &= all(pp.grad.dtype == torch.float32 for pp in ps)
    # the same exchange in the three pieces a CUDA-graph replayed step uses (pack | eager all-reduce | unpack), with
    # the early per-slice collectives deferred: identical averaged gradients
    for bdt in (None, torch.bfloat16):
        bucket = l2d.GradBucket(ps, bdt)
        bucket.defer_collectives = True
        bucket.begin()
        for pp in ps:
            pp.grad = None
        bucket.view(ps[0]).copy_(torch.full((4, 3), float(rank + 1)))
        bucket.reduce_async(ps[0])                                       # recorded only: no collective yet
        ok_bucket &= len(bucket._works) == 0
        ps[1].grad = torch.arange(7, dtype=torch.float32) * (rank + 1)
        ps[2].grad = torch.tensor(2.0 * (rank + 1))
        if rank == 0:
            ps[3].grad = torch.ones(2, 5)
        bucket.pack()
        bucket.allreduce_flat()
        bucket.unpack()
        ok_bucket &= bucket.last['calls'] == 1 and bucket.last['early_calls'] == 0
        ok_bucket &= torch.allclose(ps[0].grad, torch.full((4, 3), mean))
        ok_bucket &= torch.allclose(ps[1].grad, torch.arange(7, dtype=torch.float32) * mean)
        ok_bucket &= torch.allclose(ps[2].grad, torch.tensor(2.0 * mean))
        ok_bucket &= torch.allclose(ps[3].grad, torch.full((2, 5), 1.0 / world))
    # broadcast_module_state: rank 0's parameters, buffers and extra tensors everywhere
    torch.manual_seed(100 + rank)
    mod = torch.nn.Sequential(torch.nn.Linear(3, 2), torch.nn.BatchNorm1d(2))
    mask = torch.full((1, 6), float(rank))
    nbc = l2d.broadcast_module_state(mod, extra=[mask])
    flat = torch.cat([t.detach().reshape(-1).float() for t in list(mod.parameters()) + list(mod.buffers())] + [mask.reshape(-1)])
    both = [torch.empty_like(flat) for _ in range(world)]
    dist.all_gather(both, flat)
    ok_bcast = nbc == 8 and all(torch.equal(both[0], o) for o in both) and float(mask.sum()) == 0.0
    q.put((rank, bool(ok_gather), bool(ok_grad and ok_bucket and ok_bcast)))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('nchains,world', [(8, 2), (7, 2), (7, 3)])
def test_world2_gloo(nchains, world):
    """world 3: ragged shards and the bucket's convert-then-scale hand-back (1/3 is not exact in bf16)"""
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, nchains, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(r, True, True) for r in range(world)]

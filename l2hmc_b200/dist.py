"""Multi-GPU plumbing for the hot path: chains are independent, so
the batch is sharded contiguously over ranks with NO data-path collective; the
only exchanges are (i) an optional all_gather of per-chain observables for
logging and (ii) for L2HMC training, one flat all-reduce of the live network /
step-size gradients (the role of DDP in the reference, trainers/pytorch/
trainer.py:246-255).  One process per GPU, `torch.distributed` over NCCL
(gloo on CPU for tests)."""
from __future__ import annotations

import os
from typing import Iterable, Optional

import torch
import torch.distributed as dist

Tensor = torch.Tensor


def env_rank() -> tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment"""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def init(backend: Optional[str] = None) -> tuple[int, int, int]:
    """init_process_group(env://) with nccl on GPU / gloo on CPU, like
    `setup_torch_distributed` (utils/dist.py:126-144,228-232)"""
    rank, world, local = env_rank()
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = 'nccl' if torch.cuda.is_available() else 'gloo'
        if backend == 'nccl':
            torch.cuda.set_device(local)
        dist.init_process_group(backend)
    return rank, world, local


def rank_seed(seed: int, rank: int, local_rank: int) -> int:
    """per-rank seed, same rule as the reference (utils/dist.py:340)"""
    return seed * (rank + 1) * (local_rank + 1)


def shard_bounds(nchains: int, rank: int, world: int) -> tuple[int, int]:
    """contiguous [lo, hi) slice of the chain axis owned by `rank`; sizes differ by
    at most one and cover 0..nchains exactly"""
    if not (0 <= rank < world):
        raise ValueError(f'rank {rank} outside world of {world}')
    base, rem = divmod(nchains, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_chains(x: Tensor, rank: int, world: int) -> Tensor:
    lo, hi = shard_bounds(x.shape[0], rank, world)
    return x[lo:hi]


def gather_chains(local: Tensor, nchains: int) -> Tensor:
    """all_gather of per-chain observables (e.g. acc[nb_local]) into the global
    chain order; handles uneven shards by padding to the largest shard"""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    world = dist.get_world_size()
    sizes = [shard_bounds(nchains, r, world) for r in range(world)]
    mx = max(hi - lo for lo, hi in sizes)
    pad = torch.zeros((mx, *local.shape[1:]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:hi - lo] for o, (lo, hi) in zip(out, sizes)], 0)


def allreduce_mean_grads(params: Iterable[torch.nn.Parameter], bucket_dtype: Optional[torch.dtype] = None) -> int:
    """ONE all-reduce over a flat buffer of the gradients that exist (parameters
    that took no part in the step -- the dead SU(3) xnet -- are skipped, which is
    what `find_unused_parameters=True` buys the reference).  Returns the number of
    elements reduced."""
    grads = [p.grad for p in params if p.grad is not None]
    if not grads or not dist.is_initialized() or dist.get_world_size() == 1:
        return sum(g.numel() for g in grads)
    dt = bucket_dtype or grads[0].dtype
    flat = torch.cat([g.reshape(-1).to(dt) for g in grads])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM)
    flat /= dist.get_world_size()
    off = 0
    for g in grads:
        n = g.numel()
        g.copy_(flat[off:off + n].view_as(g))
        off += n
    return off

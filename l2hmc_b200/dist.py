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


def bind_rank_to_cores(local_rank: int, local_world: int) -> int:
    """Give each rank of a node its own slice of the host cores this process may use (and, where the GPU's PCI
    device reports a NUMA node with cores in that set, a slice of THAT node's cores), before it allocates pinned
    host memory: first-touch then places the staging buffers next to the GPU and the ranks' copy / launch threads
    stop migrating over each other's cores.  A single rank keeps every core.  Returns the number of cores bound."""
    try:
        allowed = sorted(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1
    if local_world <= 1 or len(allowed) < 2 * local_world:
        return len(allowed)
    cores = allowed
    try:
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f'{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0'
        node = int(open(f'/sys/bus/pci/devices/{bdf}/numa_node').read())
        if node >= 0:
            lst = open(f'/sys/devices/system/node/node{node}/cpulist').read().strip()
            on_node = set()
            for part in lst.split(','):
                lo, _, hi = part.partition('-')
                on_node.update(range(int(lo), int(hi or lo) + 1))
            near = [c for c in allowed if c in on_node]
            if len(near) >= 2 and len(near) < len(allowed):
                # ranks whose GPUs share this node split its cores among themselves
                peers = []
                for r in range(local_world):
                    q = torch.cuda.get_device_properties(r)
                    b2 = f'{q.pci_domain_id:04x}:{q.pci_bus_id:02x}:{q.pci_device_id:02x}.0'
                    try:
                        if int(open(f'/sys/bus/pci/devices/{b2}/numa_node').read()) == node:
                            peers.append(r)
                    except OSError:
                        pass
                if local_rank in peers and len(near) >= 2 * len(peers):
                    k = len(near) // len(peers)
                    j = peers.index(local_rank)
                    mine = near[j * k:(j + 1) * k]
                    os.sched_setaffinity(0, mine)
                    return len(mine)
    except Exception:
        pass
    k = len(cores) // local_world
    mine = cores[local_rank * k:(local_rank + 1) * k]
    try:
        os.sched_setaffinity(0, mine)
    except OSError:
        return len(allowed)
    return len(mine)


def broadcast_module_state(module: torch.nn.Module, extra: Iterable[Tensor] = (), src: int = 0) -> int:
    """What wrapping the model in DDP does at construction in the reference (trainer.py:246-255): every rank
    starts from rank `src`'s parameters and buffers.  `extra`: further tensors that must agree across ranks (the
    numpy-built leapfrog masks).  Returns the number of tensors broadcast (0 in a single process)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0
    n = 0
    with torch.no_grad():
        for t in list(module.parameters()) + list(module.buffers()) + list(extra):
            if isinstance(t, torch.nn.parameter.UninitializedParameter):
                continue
            dist.broadcast(t.data if isinstance(t, torch.nn.Parameter) else t, src)
            n += 1
    return n


class GradBucket:
    """The gradient exchange of L2HMC training (the role of DDP's buckets, trainer.py:246-255): ONE flat buffer in
    `dtype` (bf16 halves the bytes on the wire) laid out over a FIXED parameter list, so every rank reduces the same
    elements whether or not a parameter received a gradient this step (a missing gradient counts as zero).

    Producers that form a large gradient late in backward -- the deferred dW GEMMs of the vnet heads, 99.9 % of the
    bytes -- write it straight into `view(p)` and call `reduce_async(p)`: the all-reduce of that slice runs on NCCL's
    stream while the next head's GEMM and the remaining small gradients are still being computed.  `finish()` copies
    in whatever is still in `.grad`, reduces the rest in as few calls as there are contiguous gaps, waits, and hands
    every parameter its averaged gradient back in `.grad` (own dtype)."""

    def __init__(self, params: Iterable[torch.nn.Parameter], dtype: Optional[torch.dtype] = None):
        self.params = [p for p in params]
        if not self.params:
            raise ValueError('GradBucket needs at least one parameter')
        self.dtype = dtype or self.params[0].dtype
        # large tensors first: their slices are reduced individually as they land, the small rest in one call
        order = sorted(range(len(self.params)), key=lambda i: -self.params[i].numel())
        self.offsets: dict = {}
        off = 0
        for i in order:
            p = self.params[i]
            self.offsets[id(p)] = (off, p.numel())
            off += (p.numel() + 7) // 8 * 8          # 16-byte aligned slices
        self.numel = off
        self.flat = torch.zeros(off, dtype=self.dtype, device=self.params[0].device)
        self._works: list = []
        self._done: list = []                         # (offset, numel) already handed to the collective
        self.last = {'elements': off, 'bytes': off * self.flat.element_size(), 'calls': 0, 'early_calls': 0}

    def active(self) -> bool:
        return dist.is_initialized() and dist.get_world_size() > 1

    def view(self, p: torch.nn.Parameter) -> Tensor:
        off, n = self.offsets[id(p)]
        return self.flat[off:off + n].view(p.shape)

    def begin(self) -> None:
        self._works, self._done = [], []
        self.flat.zero_()

    def reduce_async(self, p: torch.nn.Parameter) -> None:
        """the gradient of `p` is complete in `view(p)`: start its all-reduce now (unless the step is being
        captured into a CUDA graph: `defer_collectives`, see pack / allreduce_flat / unpack)"""
        off, n = self.offsets[id(p)]
        self._done.append((off, n))
        if self.active() and not self.defer_collectives:
            self._works.append(dist.all_reduce(self.flat[off:off + n], op=dist.ReduceOp.SUM, async_op=True))

    # ---- the same exchange in three pieces, for a training step replayed from CUDA graphs: `pack` closes the graph
    # of forward + backward, `allreduce_flat` runs eagerly between the two graphs (a captured NCCL call replays but
    # hangs at process-group teardown), `unpack` opens the graph of clipping + Adam
    defer_collectives = False

    def pack(self) -> None:
        """everything that is still only in `.grad` goes into the flat buffer"""
        done = {o for o, _ in self._done}
        with torch.no_grad():
            for p in self.params:
                off, n = self.offsets[id(p)]
                if off not in done and p.grad is not None:
                    self.flat[off:off + n].view(p.shape).copy_(p.grad)

    def allreduce_flat(self) -> None:
        if self.active():
            dist.all_reduce(self.flat, op=dist.ReduceOp.SUM)
        self.last = {'elements': self.numel, 'bytes': self.numel * self.flat.element_size(), 'calls': 1,
                     'early_calls': 0}

    def unpack(self) -> None:
        """every parameter gets the averaged gradient back in `.grad`"""
        self._hand_back()

    def _hand_back(self) -> None:
        """every parameter of the bucket gets the averaged gradient in `.grad`, own dtype (zero where no rank produced
        one: a decision taken per rank from `p.grad is None` could differ between ranks).  ONE pass per parameter when
        the scale is a power of two (1, 2, 4, 8 ranks): a product by 2^-k is exact in the bucket's dtype, so the
        conversion to the gradient's dtype can ride on the same kernel; otherwise convert first, then scale."""
        world = dist.get_world_size() if self.active() else 1
        scale = 1.0 / world
        exact = world & (world - 1) == 0
        with torch.no_grad():
            for p in self.params:
                off, n = self.offsets[id(p)]
                src = self.flat[off:off + n].view(p.shape)
                if p.grad is None:
                    p.grad = torch.empty_like(p)
                if src.dtype == p.grad.dtype or exact:
                    torch.mul(src, scale, out=p.grad)
                else:
                    p.grad.copy_(src).mul_(scale)

    def finish(self) -> int:
        """returns the number of all-reduce calls of this step"""
        done = {o for o, _ in self._done}
        with torch.no_grad():
            for p in self.params:
                off, n = self.offsets[id(p)]
                if off not in done and p.grad is not None:
                    self.flat[off:off + n].view(p.shape).copy_(p.grad)
        early = len(self._works)
        if self.active():
            # complement of the slices already reduced, as maximal contiguous runs
            gaps, cur = [], 0
            for off, n in sorted(self._done):
                if off > cur:
                    gaps.append((cur, off))
                cur = max(cur, (off + n + 7) // 8 * 8)
            if cur < self.numel:
                gaps.append((cur, self.numel))
            for lo, hi in gaps:
                self._works.append(dist.all_reduce(self.flat[lo:hi], op=dist.ReduceOp.SUM, async_op=True))
            for w in self._works:
                w.wait()                              # stream-level wait: the current stream orders after NCCL's
        self._hand_back()
        self.last = {'elements': self.numel, 'bytes': self.numel * self.flat.element_size(),
                     'calls': len(self._works), 'early_calls': early}
        n_calls = len(self._works)
        self._works = []
        return n_calls


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

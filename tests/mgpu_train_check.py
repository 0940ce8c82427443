"""Multi-GPU check (run under torchrun on >= 2 GPUs; not collected by pytest):
L2HMC SU(3) train_step with chains sharded over ranks + ONE flat gradient all-reduce
gives every rank the same parameters, equal to the single-process full-batch step.
    python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 tests/mgpu_train_check.py"""
import os
import sys
from pathlib import Path

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from l2hmc_b200 import dist as l2d  # noqa: E402
from tests._helpers import _su3_trainer  # noqa: E402


def main():
    rank, world, local = l2d.init('nccl')
    torch.cuda.set_device(local)
    torch.set_default_dtype(torch.float64)
    nb = 4 * world

    def fresh():
        torch.manual_seed(5)
        np.random.seed(5)
        return _su3_trainer(nb=nb // world)

    # identical initial weights / masks on every rank (same seeds), global batch built identically
    tr, lat = fresh()
    torch.manual_seed(11)
    from l2hmc_b200 import ops
    xg = ops.su3_project(torch.complex(torch.randn(nb, 4, 4, 4, 4, 4, 3, 3, device='cuda'),
                                       torch.randn(nb, 4, 4, 4, 4, 4, 3, 3, device='cuda')))
    vg = ops.su3_rand_momentum(nb, [4, 4, 4, 4], 77, 0, torch.device('cuda', local))
    beta = torch.tensor(6.0)

    def grads_of(trainer, x, v):
        trainer.dynamics.train()
        trainer.optimizer.zero_grad(set_to_none=True)
        from l2hmc_b200.dynamics.pytorch.dynamics import State
        sp, met = trainer.dynamics.transition_kernel_fb(State(x, v, beta))
        loss = trainer.loss_fn(x_init=x, x_prop=sp.x, acc=met['acc'])
        loss.backward()
        return loss

    lo, hi = l2d.shard_bounds(nb, rank, world)
    grads_of(tr, xg[lo:hi].contiguous(), vg[lo:hi].contiguous())
    n = l2d.allreduce_mean_grads([p for p in tr.dynamics.parameters() if p.requires_grad])
    sharded = torch.cat([p.grad.reshape(-1) for p in tr.dynamics.parameters() if p.grad is not None])
    # reference: the whole batch in one process (same weights)
    tr2, _ = fresh()
    tr2.dynamics.config.nchains = nb
    grads_of(tr2, xg, vg)
    full = torch.cat([p.grad.reshape(-1) for p in tr2.dynamics.parameters() if p.grad is not None])
    # mean over ranks of per-shard mean losses == full-batch mean loss (equal shard sizes)
    err = float((sharded - full).abs().max() / full.abs().max())
    gathered = [torch.empty_like(sharded) for _ in range(world)]
    dist.all_gather(gathered, sharded)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    if rank == 0:
        print(f'world={world} reduced_elems={n} rel_err_vs_full_batch={err:.3e} identical_across_ranks={same}')
    assert same and err < 1e-9, (same, err)

    # ---- the Trainer path itself: ranks built from DIFFERENT seeds must start from rank 0's weights and masks
    # (broadcast at construction, the DDP wrap of trainer.py:246-255) and stay identical through train steps whose
    # gradients travel through the flat bucket (bf16 nets under autocast: the deferred head GEMMs write into it and
    # start their all-reduce early; float64 nets: everything goes through finish())
    for autocast, bdt in ((torch.bfloat16, torch.bfloat16), (None, None)):
        torch.set_default_dtype(torch.float32 if autocast is not None else torch.float64)
        torch.manual_seed(100 + rank)
        np.random.seed(100 + rank)
        tr3, lat3 = _su3_trainer(nb=2, units=(16,), autocast=autocast)
        tr3.grad_bucket_dtype = bdt
        assert tr3._bucket is not None
        if bdt is not None:
            tr3._bucket = l2d.GradBucket(tr3._bucket.params, bdt)
        torch.manual_seed(200 + rank)                 # different chains on every rank
        x3 = lat3.random().to(torch.complex128)
        before = torch.cat([p.detach().reshape(-1).double() for p in tr3.dynamics.parameters()])
        for _ in range(2):
            x3n, m3 = tr3.train_step((x3, beta))
        after = torch.cat([p.detach().reshape(-1).double() for p in tr3.dynamics.parameters()]
                          + [m.reshape(-1).double() for m in tr3.dynamics.masks])
        allp = [torch.empty_like(after) for _ in range(world)]
        dist.all_gather(allp, after)
        same3 = all(torch.equal(allp[0], a) for a in allp)
        moved = float((after[:before.numel()] - before).abs().max())
        info = tr3.allreduce_info()
        if rank == 0:
            print(f'trainer path autocast={autocast}: identical_across_ranks={same3} max|dp|={moved:.3e} exchange={info}')
        assert same3 and moved > 0 and info['calls'] >= 1 and torch.isfinite(m3['loss'])
        if autocast is not None:
            # the three head matrices and the two input-layer matrices went out from inside backward (deferred dW GEMMs)
            assert info['early_calls'] == 5, info

    # ---- the same training from CUDA graphs (graph A: forward + backward + pack, eager all-reduce of the flat
    # bucket, graph B: unpack + clip + Adam) against the eager multi-rank steps: same weights on every rank, and the
    # same trajectory of the parameters as the eager exchange up to bf16 rounding of the summation order
    torch.set_default_dtype(torch.float32)
    finals = {}
    for graphs in (False, True):
        torch.manual_seed(300)
        np.random.seed(300)
        trg, latg = _su3_trainer(nb=2, units=(16,), autocast=torch.bfloat16)
        trg.cuda_graphs = graphs
        trg.optimizer = torch.optim.Adam([p for p in trg.dynamics.parameters() if p.requires_grad], lr=1e-3,
                                         capturable=graphs, fused=True)
        trg._bucket = l2d.GradBucket(trg._bucket.params, torch.bfloat16)
        torch.manual_seed(400 + rank)
        xg = latg.random().to(torch.complex128)
        torch.manual_seed(500 + rank)                 # same momenta / accept draws in both runs
        for _ in range(5):
            _, mg = trg.train_step((xg, beta))
        assert torch.isfinite(mg['loss'])
        pf = torch.cat([p.detach().reshape(-1).double() for p in trg.dynamics.parameters()])
        allg = [torch.empty_like(pf) for _ in range(world)]
        dist.all_gather(allg, pf)
        assert all(torch.equal(allg[0], a) for a in allg), 'ranks diverged'
        finals[graphs] = pf
    if rank == 0:
        print(f'graphed multi-rank training: exchange={trg.allreduce_info()} '
              f'finite={bool(torch.isfinite(finals[True]).all())} moved={float((finals[True] - finals[False]).abs().max()):.3e}')
    assert bool(torch.isfinite(finals[True]).all())
    torch.set_default_dtype(torch.float64)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()

"""Minimal re-host of the three `Trainer` step functions that enter the hot path
(`trainers/pytorch/trainer.py:904-956,1266-1367`): `hmc_step`, `eval_step`,
`train_step`, with the reference's call contract (x is projected with
`g.compat_proj` at the top of every step; the loss sees `mc_states.proposed.x`;
the trainer detaches `x_out`).  Everything else the reference's Trainer does
(W&B/Aim, rich tables, checkpoints, schedules) is orchestration and out of scope
(SURVEY section 2); multi-GPU training replaces DDP by ONE flat all-reduce of the
gradients that exist (l2hmc_b200/dist.py)."""
from __future__ import annotations

from typing import Optional

import torch

from ... import _lib
from ... import autograd as ag
from ... import dist as l2dist
from ...configs import LossConfig
from ...dynamics.pytorch.dynamics import Dynamics
from ...loss.pytorch.loss import LatticeLoss
from ...network.pytorch.network import bump_weights_generation, new_step_token, weights_generation

Tensor = torch.Tensor


class Trainer:
    def __init__(self, dynamics: Dynamics, loss_config: Optional[LossConfig] = None, lr: float = 1e-3,
                 clip_val: float = 0.0, autocast_dtype: Optional[torch.dtype] = None,
                 grad_bucket_dtype: Optional[torch.dtype] = None, cuda_graphs: bool = False):
        """cuda_graphs: capture each step function (one graph per input shape / beta / step
        arguments) on first use and replay it afterwards -- the whole step, including backward
        and Adam for `train_step`, becomes ONE graph launch.  Step sizes and the
        SU(3) momentum RNG counter live on the device, so replays stay correct while parameters
        train.  Needs `merge_directions` (the default).  With more than one rank `train_step` replays as two
        graphs around one eager NCCL all-reduce of the flat gradient bucket (`_graphed_train_multi`)."""
        self.dynamics = dynamics
        self.lattice = dynamics.lattice
        self.g = dynamics.g
        self.loss_fn = LatticeLoss(self.lattice, loss_config or LossConfig())
        params = [p for p in dynamics.parameters() if p.requires_grad]
        self.cuda_graphs = bool(cuda_graphs)
        # fused: one multi-tensor kernel over the 180 M vnet parameters instead of the foreach chain
        self.optimizer = torch.optim.Adam(params, lr=lr, capturable=self.cuda_graphs, fused=True)
        self.clip_val = clip_val
        self.autocast_dtype = autocast_dtype
        self.grad_bucket_dtype = grad_bucket_dtype
        self._graphs: dict = {}
        # kernels of libl2b recorded into each captured step ('hmc' / 'eval' / 'train'): what ONE replay launches
        # (the host-side launch counter does not move on a replay)
        self.graph_launches: dict = {}
        self._eager = False            # True while warming up / capturing: step functions run their eager body
        # more than one rank: start from rank 0's weights, buffers and leapfrog masks (what wrapping the model in
        # DDP does in the reference, trainer.py:246-255), and exchange gradients through one flat bucket over the
        # FIXED list of trainable parameters
        self._bucket = None
        if l2dist.dist.is_available() and l2dist.dist.is_initialized() and l2dist.dist.get_world_size() > 1:
            l2dist.broadcast_module_state(dynamics, extra=getattr(dynamics, 'masks', ()))
            bump_weights_generation()      # in-place on .data: no `_version` moved
            if hasattr(dynamics, '_planar_cache'):
                dynamics._planar_cache = None
            self._bucket = l2dist.GradBucket(params, grad_bucket_dtype)

    def allreduce_info(self) -> Optional[dict]:
        """size / dtype / number of collectives of the last gradient exchange (None in a single process)"""
        if self._bucket is None:
            return None
        return dict(self._bucket.last, dtype=str(self._bucket.dtype).replace('torch.', ''),
                    world=l2dist.dist.get_world_size())

    # ------------------------------------------------------------ CUDA graphs
    def _canon(self, x: Tensor) -> Tensor:
        """on the device, in the lattice shape (the flattened x_out of the previous step maps to the same graph)"""
        x = x.to(self.dynamics.device)
        return x.reshape(x.shape[0], *self.dynamics.xshape[1:])

    def _weights_version(self) -> tuple:
        """changes whenever the parameters may have: Python-side mutation (`_version`) or a training step,
        eager or replayed from a graph (the generation counter `train_step` bumps)"""
        return (sum(p._version for p in self.dynamics.parameters()), weights_generation())

    def _graphed(self, key, fn, x: Tensor, train: bool = False):
        """run `fn(static_x) -> (x_out, metrics)` through a CUDA graph keyed by `key`"""
        ent = self._graphs.get(key)
        if ent is None:
            if not self.dynamics.config.merge_directions:
                raise RuntimeError('cuda_graphs needs merge_directions=True (the direction coin is drawn on the host)')
            static_x = x.detach().clone()
            if getattr(self.g, '_name', None) == 'SU3':
                from ...group.su3.pytorch import group as su3group
                su3group._graph_counter(static_x.device)
            self._eager = True
            try:
                side = torch.cuda.Stream(device=static_x.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):           # warm-up off the capture: lazy allocations, workspaces
                    for _ in range(3):                  # (for train_step these are three ordinary training steps)
                        fn(static_x)
                torch.cuda.current_stream().wait_stream(side)
                ag._BAD_FLAGS.clear()
                if train:
                    self.optimizer.zero_grad(set_to_none=True)
                graph = torch.cuda.CUDAGraph()
                n0 = _lib.launch_count()
                # thread_local: the NCCL watchdog thread may touch the CUDA runtime while we capture
                with torch.cuda.graph(graph, capture_error_mode='thread_local'):
                    out = fn(static_x)
                self.graph_launches[key[0]] = _lib.launch_count() - n0
            finally:
                self._eager = False
            flags = list(ag._BAD_FLAGS)                 # matrix-exp adjoint range flags: static, re-read per replay
            ag._BAD_FLAGS.clear()
            ent = (graph, static_x, out, flags)
            if key[0] == 'eval':
                # an eval graph bakes in the weights-derived caches of its capture (bf16 head image, step sizes):
                # graphs of older weights can never be replayed again -- drop them instead of growing the pool
                for old in [k for k in self._graphs if k[0] == 'eval' and k[1:4] == key[1:4] and k[5] == key[5] and k != key]:
                    del self._graphs[old]
            self._graphs[key] = ent
        graph, static_x, out, flags = ent
        static_x.copy_(x)
        graph.replay()
        if flags and int(torch.stack([f.reshape(()) for f in flags]).sum()) != 0:
            raise ag.ops.L2BError('matrix-exponential adjoint: ||eps p||_F > 3 (outside the series\' validated range)')
        xo, metrics = out
        return xo.clone(), {k: (v.clone() if isinstance(v, Tensor) else v) for k, v in metrics.items()}

    def _x(self, x: Tensor) -> Tensor:
        return self.g.compat_proj(x.reshape(x.shape[0], *self.dynamics.xshape[1:]))

    @torch.no_grad()
    def hmc_step(self, inputs, eps: Optional[float] = None, nleapfrog: Optional[int] = None):
        """trainer.py:904-929"""
        if self.cuda_graphs and not self._eager:
            xi, beta = inputs
            xi = self._canon(xi)
            key = ('hmc', tuple(xi.shape), xi.dtype, float(beta), eps, nleapfrog)
            return self._graphed(key, lambda xs: self.hmc_step((xs, float(beta)), eps=eps, nleapfrog=nleapfrog), xi)
        xi, beta = inputs
        xi = self._x(xi.to(self.dynamics.device))
        xo, metrics = self.dynamics.apply_transition_hmc((xi, beta), eps=eps, nleapfrog=nleapfrog)
        xp = metrics.pop('mc_states').proposed.x
        loss = self.loss_fn(x_init=xi, x_prop=xp, acc=metrics['acc'])
        if self.dynamics.config.verbose:      # trainer.py:922-924: plaqs, charges, dQint / dQsin join the metrics
            metrics.update(self.loss_fn.lattice_metrics(xinit=xi, xout=xo))
        metrics['loss'] = loss
        return xo.detach(), metrics

    @torch.no_grad()
    def eval_step(self, inputs):
        """trainer.py:931-956"""
        if self.cuda_graphs and not self._eager:
            xi, beta = inputs
            xi = self._canon(xi)
            key = ('eval', tuple(xi.shape), xi.dtype, float(beta), self._weights_version(),
                   torch.is_autocast_enabled('cuda'))
            return self._graphed(key, lambda xs: self.eval_step((xs, float(beta))), xi)
        self.dynamics.eval()
        xi, beta = inputs
        xi = self._x(xi.to(self.dynamics.device))
        xo, metrics = self.dynamics((xi, beta))
        xp = metrics.pop('mc_states').proposed.x
        metrics['loss'] = self.loss_fn(x_init=xi, x_prop=xp, acc=metrics['acc'])
        if self.dynamics.config.verbose:      # trainer.py:948-950
            metrics.update(self.loss_fn.lattice_metrics(xinit=xi, xout=xo))
        return xo.detach(), metrics

    def warmup(self, beta, nsteps: int = 100, tol: float = 1e-5, x: Optional[Tensor] = None,
               nchains: Optional[int] = None) -> Tensor:
        """Thermalise configurations with accept / reject HMC steps (trainer.py:1699-1744).  For U(1) with verbose
        metrics the loop stops early once the plaquette agrees with the exact I1(beta) / I0(beta) to `tol` (summed
        over chains), as upstream; SU(3) runs all `nsteps`."""
        self.dynamics.eval()
        if x is None:
            x = self.dynamics.lattice.random().to(self.dynamics.device)
        if nchains is not None:
            x = x[:nchains]
        if not isinstance(beta, Tensor):
            beta = torch.tensor(float(beta))
        pexact = None
        if self.dynamics.config.group == 'U1':
            from ...lattice.u1.pytorch.lattice import plaq_exact
            pexact = plaq_exact(beta).to(self.dynamics.device)
        for _ in range(nsteps):
            x, metrics = self.hmc_step((x, beta))
            plaqs = metrics.get('plaqs', None)
            if plaqs is not None and pexact is not None and float((plaqs - pexact).abs().sum()) < tol:
                return x
        self.dynamics.train()
        return x

    def train_step(self, inputs):
        """forward, loss, backward, (all-reduce), clip, Adam   (trainer.py:1266-1367)"""
        if self.cuda_graphs and not self._eager:
            xi, beta = inputs
            xi = self._canon(xi)
            key = ('train', tuple(xi.shape), xi.dtype, float(beta))
            if self._bucket is not None and self._bucket.active():
                out = self._graphed_train_multi(key, xi, float(beta))
            else:
                out = self._graphed(key, lambda xs: self.train_step((xs, float(beta))), xi, train=True)
            bump_weights_generation()     # the replay moved the weights without Python touching a parameter
            return out
        xo, metrics = self._forward_backward(inputs)
        # DDP's job in the reference (trainer.py:246-255): mean of the gradients over ranks.  The large matrices
        # (99.9 % of the bytes) were written into the bucket by their deferred GEMMs and are already on the wire;
        # finish() adds the small rest, waits, and leaves the averaged gradients in .grad
        if self._bucket is not None:
            self._bucket.finish()
        self._apply_gradients()
        return xo, metrics

    def _forward_backward(self, inputs):
        """first half of a training step: forward, loss, backward (gradients in `.grad` / the exchange bucket)"""
        self.dynamics.train()
        new_step_token()        # weight-derived caches built during a capture live for exactly this step
        xi, beta = inputs
        with torch.no_grad():
            xi = self._x(xi.to(self.dynamics.device))
        capturing = torch.cuda.is_current_stream_capturing()
        if not capturing:       # in a captured step the gradients are static buffers the replay overwrites
            self.optimizer.zero_grad(set_to_none=True)
        if self.autocast_dtype is not None:
            with torch.autocast('cuda', dtype=self.autocast_dtype):
                xo, metrics = self.dynamics((xi, beta))
        else:
            xo, metrics = self.dynamics((xi, beta))
        xp = metrics.pop('mc_states').proposed.x
        loss = self.loss_fn(x_init=xi, x_prop=xp, acc=metrics['acc'])
        ag.DEFER_HEAD_GRADS = True        # one dW GEMM per weight matrix per step instead of one per v-update
        if self._bucket is not None:
            self._bucket.begin()
            ag.HEAD_GRAD_SINK = self._bucket
        try:
            loss.backward()
        finally:
            ag.DEFER_HEAD_GRADS = False
            ag.HEAD_GRAD_SINK = None
        metrics['loss'] = loss.detach()
        return xo.detach(), metrics

    def _apply_gradients(self) -> None:
        """second half: matrix-exponential adjoint range check, clipping, Adam"""
        if not torch.cuda.is_current_stream_capturing():
            ag.check_exp_adjoint_flags()  # one device read per step
        if self.clip_val > 0:
            torch.nn.utils.clip_grad_norm_(self.optimizer.param_groups[0]['params'], self.clip_val)
        self.optimizer.step()
        bump_weights_generation()

    def _graphed_train_multi(self, key, x: Tensor, beta: float):
        """Multi-rank training from CUDA graphs: graph A = forward + backward + packing the gradient bucket, ONE
        eager NCCL all-reduce of the flat bf16 bucket, graph B = unpacking + clipping + Adam.  (The eager multi-rank
        step is host-bound once eight ranks share the node's cores: 48 ms against 25 ms on one GPU.  The per-slice
        early all-reduces of the eager path are given up here: one 362 MB all-reduce over NVLink costs ~1.5 ms.)"""
        ent = self._graphs.get(key)
        bucket = self._bucket
        if ent is None:
            if not self.dynamics.config.merge_directions:
                raise RuntimeError('cuda_graphs needs merge_directions=True (the direction coin is drawn on the host)')
            static_x = x.detach().clone()
            if getattr(self.g, '_name', None) == 'SU3':
                from ...group.su3.pytorch import group as su3group
                su3group._graph_counter(static_x.device)
            self._eager = True
            try:
                side = torch.cuda.Stream(device=static_x.device)
                side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(side):           # three ordinary (eager) training steps: lazy allocations
                    for _ in range(3):
                        self.train_step((static_x, beta))
                torch.cuda.current_stream().wait_stream(side)
                torch.cuda.synchronize()
                ag._BAD_FLAGS.clear()
                self.optimizer.zero_grad(set_to_none=True)
                bucket.defer_collectives = True
                graph_a, graph_b = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
                n0 = _lib.launch_count()
                with torch.cuda.graph(graph_a, capture_error_mode='thread_local'):
                    out = self._forward_backward((static_x, beta))
                    bucket.pack()
                flags = list(ag._BAD_FLAGS)
                ag._BAD_FLAGS.clear()
                bucket.allreduce_flat()                  # the capture pass is a real step on every rank
                with torch.cuda.graph(graph_b, pool=graph_a.pool(), capture_error_mode='thread_local'):
                    bucket.unpack()
                    self._apply_gradients()
                self.graph_launches[key[0]] = _lib.launch_count() - n0
            finally:
                bucket.defer_collectives = False
                self._eager = False
            ent = (graph_a, graph_b, static_x, out, flags)
            self._graphs[key] = ent
        graph_a, graph_b, static_x, out, flags = ent
        static_x.copy_(x)
        graph_a.replay()
        bucket.allreduce_flat()
        graph_b.replay()
        if flags and int(torch.stack([f.reshape(()) for f in flags]).sum()) != 0:
            raise ag.ops.L2BError('matrix-exponential adjoint: ||eps p||_F > 3 (outside the series\' validated range)')
        xo, metrics = out
        return xo.clone(), {k: (v.clone() if isinstance(v, Tensor) else v) for k, v in metrics.items()}

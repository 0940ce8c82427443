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

from ... import autograd as ag
from ... import dist as l2dist
from ...configs import LossConfig
from ...dynamics.pytorch.dynamics import Dynamics
from ...loss.pytorch.loss import LatticeLoss

Tensor = torch.Tensor


class Trainer:
    def __init__(self, dynamics: Dynamics, loss_config: Optional[LossConfig] = None, lr: float = 1e-3,
                 clip_val: float = 0.0, autocast_dtype: Optional[torch.dtype] = None,
                 grad_bucket_dtype: Optional[torch.dtype] = None):
        self.dynamics = dynamics
        self.lattice = dynamics.lattice
        self.g = dynamics.g
        self.loss_fn = LatticeLoss(self.lattice, loss_config or LossConfig())
        params = [p for p in dynamics.parameters() if p.requires_grad]
        self.optimizer = torch.optim.Adam(params, lr=lr)
        self.clip_val = clip_val
        self.autocast_dtype = autocast_dtype
        self.grad_bucket_dtype = grad_bucket_dtype

    def _x(self, x: Tensor) -> Tensor:
        return self.g.compat_proj(x.reshape(x.shape[0], *self.dynamics.xshape[1:]))

    @torch.no_grad()
    def hmc_step(self, inputs, eps: Optional[float] = None, nleapfrog: Optional[int] = None):
        """trainer.py:904-929"""
        xi, beta = inputs
        xi = self._x(xi.to(self.dynamics.device))
        xo, metrics = self.dynamics.apply_transition_hmc((xi, beta), eps=eps, nleapfrog=nleapfrog)
        xp = metrics.pop('mc_states').proposed.x
        loss = self.loss_fn(x_init=xi, x_prop=xp, acc=metrics['acc'])
        metrics['loss'] = loss
        return xo.detach(), metrics

    @torch.no_grad()
    def eval_step(self, inputs):
        """trainer.py:931-956"""
        self.dynamics.eval()
        xi, beta = inputs
        xi = self._x(xi.to(self.dynamics.device))
        xo, metrics = self.dynamics((xi, beta))
        xp = metrics.pop('mc_states').proposed.x
        metrics['loss'] = self.loss_fn(x_init=xi, x_prop=xp, acc=metrics['acc'])
        return xo.detach(), metrics

    def train_step(self, inputs):
        """forward, loss, backward, (all-reduce), clip, Adam   (trainer.py:1266-1367)"""
        self.dynamics.train()
        xi, beta = inputs
        with torch.no_grad():
            xi = self._x(xi.to(self.dynamics.device))
        self.optimizer.zero_grad(set_to_none=True)
        if self.autocast_dtype is not None:
            with torch.autocast('cuda', dtype=self.autocast_dtype):
                xo, metrics = self.dynamics((xi, beta))
        else:
            xo, metrics = self.dynamics((xi, beta))
        xp = metrics.pop('mc_states').proposed.x
        loss = self.loss_fn(x_init=xi, x_prop=xp, acc=metrics['acc'])
        loss.backward()
        ag.check_exp_adjoint_flags()      # one device read per step (matrix-exp adjoint range check)
        # DDP's job in the reference (trainer.py:246-255): mean of the gradients over ranks;
        # parameters without a gradient (the unused SU(3) xnet) are not communicated.
        l2dist.allreduce_mean_grads(self.optimizer.param_groups[0]['params'], self.grad_bucket_dtype)
        if self.clip_val > 0:
            torch.nn.utils.clip_grad_norm_(self.optimizer.param_groups[0]['params'], self.clip_val)
        self.optimizer.step()
        metrics['loss'] = loss.detach()
        return xo.detach(), metrics

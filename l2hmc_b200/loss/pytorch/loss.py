"""`LatticeLoss` with the reference's surface (`loss/pytorch/loss.py:20-210`):
ESJD-style losses on (x_init, x_prop, acc) built from Wilson loops.  The Wilson
loops come from libl2b (differentiable, l2hmc_b200/autograd.py); the handful of
[nb]-sized reductions on top of them are plain torch."""
from __future__ import annotations

from typing import Optional

import torch

from ...configs import LossConfig
from ...lattice.su3.pytorch.lattice import LatticeSU3
from ...lattice.u1.pytorch.lattice import LatticeU1

Tensor = torch.Tensor


class LatticeLoss:
    def __init__(self, lattice: LatticeU1 | LatticeSU3, loss_config: LossConfig):
        self.lattice = lattice
        self.config = loss_config
        self.xshape = self.lattice._shape
        self.plaq_weight = torch.tensor(self.config.plaq_weight, dtype=torch.float)
        self.charge_weight = torch.tensor(self.config.charge_weight, dtype=torch.float)
        self.rmse_weight = torch.tensor(self.config.rmse_weight, dtype=torch.float)
        if not isinstance(lattice, (LatticeU1, LatticeSU3)):
            raise ValueError(f'Unexpected lattice: {lattice}')
        self.g = lattice.g
        self._dev_weights: dict = {}

    def _w(self, w: Tensor, like: Tensor) -> Tensor:
        """loss weight on `like`'s device, copied once (an H2D copy per call would synchronise and
        cannot be captured in a CUDA graph)"""
        key = (id(w), like.device)
        if key not in self._dev_weights:
            self._dev_weights[key] = w.to(like.device)
        return self._dev_weights[key]

    def __call__(self, x_init: Tensor, x_prop: Tensor, acc: Tensor) -> Tensor:
        return self.calc_loss(x_init=x_init, x_prop=x_prop, acc=acc)

    @staticmethod
    def mixed_loss(loss: Tensor, weight: Tensor) -> Tensor:
        return (weight / loss) - (loss / weight)

    def _plaq_loss(self, w1: Tensor, w2: Tensor, acc: Tensor, use_mixed_loss: Optional[bool] = None) -> Tensor:
        p1 = w1.real.sum(list(range(2, len(w1.shape))))
        p2 = w2.real.sum(list(range(2, len(w2.shape))))
        ploss = acc * (p2 - p1) ** 2
        if use_mixed_loss:
            ploss = ploss + 1e-4
            return self.mixed_loss(ploss, self._w(self.plaq_weight, ploss)).mean()
        return (-ploss / self._w(self.plaq_weight, ploss)).mean()

    def _charge_loss(self, w1: Tensor, w2: Tensor, acc: Tensor, use_mixed_loss: Optional[bool] = None) -> Tensor:
        q1 = self.lattice._sin_charges(wloops=w1)
        q2 = self.lattice._sin_charges(wloops=w2)
        qloss = acc * (q2 - q1) ** 2
        use_mixed = self.config.use_mixed_loss if use_mixed_loss is None else use_mixed_loss
        if use_mixed:
            qloss = qloss + 1e-4
            return self.mixed_loss(qloss, self._w(self.charge_weight, qloss)).mean()
        return (-qloss / self._w(self.charge_weight, qloss)).mean()

    def lattice_metrics(self, xinit: Tensor, xout: Optional[Tensor] = None) -> dict[str, Tensor]:
        metrics = self.lattice.calc_metrics(x=xinit)
        if xout is not None:
            wloops = self.lattice.wilson_loops(x=xout)
            qint = self.lattice._int_charges(wloops=wloops)
            qsin = self.lattice._sin_charges(wloops=wloops)
            metrics.update({'dQint': (qint - metrics['intQ']).abs(), 'dQsin': (qsin - metrics['sinQ']).abs()})
        return metrics

    def plaq_loss(self, x_init, x_prop, acc, use_mixed_loss: Optional[bool] = None) -> Tensor:
        return self._plaq_loss(self.lattice.wilson_loops(x=x_init), self.lattice.wilson_loops(x=x_prop), acc,
                               use_mixed_loss)

    def charge_loss(self, x_init, x_prop, acc, use_mixed_loss: Optional[bool] = None) -> Tensor:
        return self._charge_loss(self.lattice.wilson_loops(x=x_init), self.lattice.wilson_loops(x=x_prop), acc,
                                 use_mixed_loss)

    def rmse_loss(self, x_init, x_prop, acc, use_mixed_loss: Optional[bool] = None) -> Tensor:
        dx = x_prop.reshape(x_init.shape) - x_init
        dx2 = (dx.real ** 2 + dx.imag ** 2).flatten(1) if dx.is_complex() else (dx ** 2).flatten(1)
        rmse_loss = acc * dx2.mean(1)
        use_mixed = self.config.use_mixed_loss if use_mixed_loss is None else use_mixed_loss
        if use_mixed:
            rmse_loss = rmse_loss + 1e-4
            return self.mixed_loss(rmse_loss, self._w(self.rmse_weight, rmse_loss)).mean()
        return (-rmse_loss / self._w(self.rmse_weight, rmse_loss)).mean()

    def general_loss(self, x_init, x_prop, acc, plaq_weight=None, charge_weight=None, use_mixed_loss=None):
        wl_init = self.lattice.wilson_loops(x=x_init)
        wl_prop = self.lattice.wilson_loops(x=x_prop)
        pw = self.plaq_weight if plaq_weight is None else plaq_weight
        qw = self.charge_weight if charge_weight is None else charge_weight
        loss = 0.0
        if pw > 0:
            loss = loss + pw * self._plaq_loss(wl_init, wl_prop, acc, use_mixed_loss)
        if qw > 0:
            loss = loss + qw * self._charge_loss(wl_init, wl_prop, acc, use_mixed_loss)
        return loss

    def calc_loss(self, x_init: Tensor, x_prop: Tensor, acc: Tensor) -> Tensor:
        """loss.py:194-210"""
        x_prop = x_prop.reshape(x_init.shape)
        wl_init = self.lattice.wilson_loops(x=x_init)
        wl_prop = self.lattice.wilson_loops(x=x_prop)
        zero = acc.new_zeros(())
        rmse = self.rmse_loss(x_init, x_prop, acc) if self.rmse_weight > 0 else zero
        plaq = self._plaq_loss(wl_init, wl_prop, acc) if self.plaq_weight > 0 else zero
        charge = self._charge_loss(wl_init, wl_prop, acc) if self.charge_weight > 0 else zero
        return plaq + charge + rmse

"""`U1Phase` group object with the reference's method surface
(`group/u1/pytorch/group.py:60-165`); x are real angles."""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from ....group.group import Group
from .... import ops
from .... import autograd as ag

PI = torch.pi
TWO_PI = 2. * torch.pi
Tensor = torch.Tensor


def _device() -> torch.device:
    if not torch.cuda.is_available():
        raise ops.L2BError('l2hmc_b200 needs a CUDA device (no CPU fallback)')
    return torch.device('cuda', torch.cuda.current_device())


def rand_unif(shape: Sequence[int], a: float, b: float, requires_grad: bool = True) -> Tensor:
    """x ~ U(a, b) (group.py:23-41), on the CUDA device"""
    rand = (a - b) * torch.rand(tuple(shape), device=_device()) + b
    return rand.clone().detach().requires_grad_(requires_grad)


def random_angle(shape: Sequence[int], requires_grad: bool = True) -> Tensor:
    """group.py:44-47"""
    return rand_unif(shape, -PI, PI, requires_grad=requires_grad)


def eyeOf(x: Tensor) -> Tensor:
    """group.py:50-57"""
    eye = torch.zeros([1] * (x.dim() - 1) + [*x.shape[-1:]], device=x.device)
    eye[-1:] = torch.eye(x.shape[-1], device=x.device)
    return eye


class U1Phase(Group):
    def __init__(self) -> None:
        super().__init__(dim=2, shape=[1], dtype=torch.get_default_dtype())

    def phase_to_coords(self, phi: Tensor) -> Tensor:
        return torch.cat([phi.cos(), phi.sin()], -1)

    def coords_to_phase(self, x: Tensor) -> Tensor:
        assert x.shape[-1] == 2
        return torch.atan2(x[..., -1], x[..., -2])

    @staticmethod
    def group_to_vec(x: Tensor) -> Tensor:
        """[cos x, sin x] concatenated on dim 1 (network input packing, group.py:86-88)"""
        return torch.cat([x.cos(), x.sin()], dim=1)

    @staticmethod
    def vec_to_group(x: Tensor) -> Tensor:
        if x.is_complex():
            return torch.atan2(x.imag, x.real)
        return torch.atan2(x[..., -1], x[..., -2])

    def exp(self, x: Tensor) -> Tensor:
        return torch.complex(x.cos(), x.sin())

    def update_gauge(self, x: Tensor, p: Tensor) -> Tensor:
        return x + p

    def mul(self, a: Tensor, b: Tensor, adjoint_a: Optional[bool] = None, adjoint_b: Optional[bool] = None) -> Tensor:
        if adjoint_a and adjoint_b:
            return -a - b
        if adjoint_a:
            return -a + b
        if adjoint_b:
            return a - b
        return a + b

    def adjoint(self, x: Tensor) -> Tensor:
        return -x

    def trace(self, x: Tensor) -> Tensor:
        return torch.cos(x)

    def diff_trace(self, x: Tensor) -> Tensor:
        return -torch.sin(x)

    def diff2trace(self, x: Tensor) -> Tensor:
        return -torch.cos(x)

    def floormod(self, x, y) -> Tensor:
        """group.py:134-135"""
        return x - torch.floor_divide(x, y) * y

    def compat_proj(self, x: Tensor) -> Tensor:
        """((x + pi) mod 2 pi) - pi   (group.py:130-131)"""
        return ag.U1CompatProj.apply(x)

    def projectTAH(self, x: Tensor) -> Tensor:
        return x

    def projectSU(self, x: Tensor) -> Tensor:
        return self.compat_proj(x)

    def random(self, shape: Sequence[int]) -> Tensor:
        return self.compat_proj(TWO_PI * torch.rand(*shape, device=_device()))

    def random_momentum(self, shape: Sequence[int]) -> Tensor:
        return torch.randn(*shape, device=_device()).reshape(shape[0], -1)

    def kinetic_energy(self, p: Tensor) -> Tensor:
        return ag.U1Kinetic.apply(p)

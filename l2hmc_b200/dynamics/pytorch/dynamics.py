"""`Dynamics`: the generalised leapfrog integrator + Metropolis-Hastings step,
same public surface as the reference's
`dynamics/pytorch/dynamics.py:113-1535`, with the per-step work done by the
sm_100a kernels of libl2b:

  plain HMC  (`apply_transition_hmc` / `transition_kernel_hmc`, :632-658,:915-954)
      one call of `l2b_{su3,u1}_hmc_trajectory` (whole trajectory; energies for
      the accept probability come fused with the first / last force pass);
  L2HMC      (`forward` -> `apply_transition_fb` / `transition_kernel_fb`, :956-1029)
      force, projectSU+su3_to_vec, v-update epilogue, masked x-update and the
      Hamiltonians are libl2b kernels; the xnet/vnet dense layers are torch.nn.

Reference quirks kept on purpose: eps -> eps/(1+eps) in
L2HMC only; element-wise float32 masks built with numpy's RNG; SU(3) x-update
`m*x + exp(eps v) @ ((1-m)*x)` with logdet 0 and no xnet call; vnet inputs go
through projectSU; `acc_mask` float32; HMC `nleapfrog` doubles when
`merge_directions`; x_out returned flattened.

Training: both L2HMC paths are differentiable end to end through hand-written adjoint
kernels (l2hmc_b200/autograd.py): U(1) force / updates / loops; SU(3) action, force,
v-update (also fused with the tcgen05 heads), masked exp(eps v) x-update
(matrix-exponential adjoint), projectSU + su3_to_vec (closed-form polar-factor adjoint),
per-site Wilson loops and the kinetic energy.  Only the dense layers in front of the
output heads and the rectangle term of the improved action (c1 != 0) use torch autograd.
"""
from __future__ import annotations

import os
from dataclasses import dataclass
from math import pi as PI
from pathlib import Path
from typing import Callable, Optional, Union

import numpy as np
import torch
from torch import nn

from ... import configs as cfgs
from ... import ops
from ... import autograd as ag
from ...group.su3.pytorch.group import SU3
from ...group.u1.pytorch.group import U1Phase
from ...lattice.su3.pytorch.lattice import LatticeSU3
from ...lattice.u1.pytorch.lattice import LatticeU1
from ...network.pytorch.network import NetworkFactory, dummy_network, weights_generation

TWO_PI = 2. * PI
Shape = Union[tuple, list]
Tensor = torch.Tensor


@dataclass
class State:
    x: Tensor
    v: Tensor
    beta: Tensor

    def __post_init__(self):
        self.nb = self.x.shape[0]
        self.xshape = self.x.shape

    def flatten(self) -> 'State':
        return State(x=self.x.flatten(1), v=self.v.flatten(1), beta=self.beta)

    def to_numpy(self):
        return {'x': self.x.detach().cpu().numpy(), 'v': self.v.detach().cpu().numpy(),
                'beta': torch.as_tensor(self.beta).detach().cpu().numpy()}


@dataclass
class MonteCarloStates:
    init: State
    proposed: State
    out: State


def rand_unif(shape: Shape, a: float, b: float, requires_grad: bool) -> Tensor:
    """uniform samples between a and b (dynamics.py:86-94)"""
    return ((a - b) * torch.rand(tuple(shape)) + b).detach().requires_grad_(requires_grad)


def random_angle(shape: Shape, requires_grad: bool = True) -> Tensor:
    """angles uniform in (-pi, pi)   (dynamics.py:97-99)"""
    return rand_unif(shape, -PI, PI, requires_grad=requires_grad)


def to_u1(x: Tensor) -> Tensor:
    """wrap to [-pi, pi)   (dynamics.py:77-79)"""
    return ((x + PI) % TWO_PI) - PI


def sigmoid(x: Tensor) -> Tensor:
    return 1. / (1. + torch.exp(-x))


class Mask:
    """m and its complement 1 - m (dynamics.py:102-110; unused by the reference's own integrator)"""

    def __init__(self, m: Tensor):
        self.m = m
        self.mb = torch.ones_like(self.m) - self.m

    def combine(self, x: Tensor, y: Tensor) -> Tensor:
        return self.m * x + self.mb * y


def _fbeta(beta) -> float:
    return float(beta.detach()) if isinstance(beta, torch.Tensor) else float(beta)


class Dynamics(nn.Module):
    def __init__(self, potential_fn: Callable, config: cfgs.DynamicsConfig,
                 network_factory: Optional[NetworkFactory] = None):
        super().__init__()
        if not torch.cuda.is_available():
            raise ops.L2BError('l2hmc_b200.Dynamics needs a CUDA device (no CPU fallback)')
        self.config = config
        self.xdim = self.config.xdim
        self.xshape = self.config.xshape
        self.potential_fn = potential_fn
        self.nlf = self.config.nleapfrog
        self.device = self._device = torch.device('cuda', torch.cuda.current_device())
        self._su3 = self.config.group.upper() == 'SU3'
        if self._su3:
            self.g = SU3()
            self.lattice = LatticeSU3(self.config.nchains, self.config.latvolume)
        else:
            self.g = U1Phase()
            self.lattice = LatticeU1(self.config.nchains, self.config.latvolume)
        self.network_factory = network_factory
        if network_factory is not None:
            self._networks_built = True
            self.networks = self._build_networks(network_factory)
            # registered twice, like the reference (dynamics.py:140-144): the
            # state_dict carries both `networks.xnet.*` and `xnet.*`
            self.xnet = self.networks['xnet']
            self.vnet = self.networks['vnet']
        else:
            self._networks_built = False
            self.xnet = dummy_network
            self.vnet = dummy_network
            self.networks = {'xnet': self.xnet, 'vnet': self.vnet}
        self.masks = [m.to(self.device) for m in self._build_masks()]
        self._dtype = torch.get_default_dtype()
        rg = (not self.config.eps_fixed)
        self.xeps = nn.ParameterList([
            nn.Parameter(torch.tensor(float(self.config.eps)), requires_grad=rg)
            for _ in range(self.config.nleapfrog)])
        self.veps = nn.ParameterList([
            nn.Parameter(torch.tensor(float(self.config.eps)), requires_grad=rg)
            for _ in range(self.config.nleapfrog)])
        self.to(self.device)

    # ------------------------------------------------------------------ build
    def _build_networks(self, network_factory: NetworkFactory) -> nn.ModuleDict:
        split = self.config.use_split_xnets
        n = self.nlf if self.config.use_separate_networks else 1
        return network_factory.build_networks(n, split, group=self.g)

    def _build_masks(self):
        """nlf random half-masks over the xdim ELEMENTS, numpy RNG
        (dynamics.py:1101-1110)"""
        masks = []
        for _ in range(self.config.nleapfrog):
            idx = np.random.permutation(np.arange(self.xdim))[:self.xdim // 2]
            mask = np.zeros((self.xdim,), dtype=np.float32)
            mask[idx] = 1.
            masks.append(torch.from_numpy(mask[None, :]))
        return masks

    def get_models(self) -> dict:
        if self.config.use_separate_networks:
            xnet, vnet = {}, {}
            for lf in range(self.config.nleapfrog):
                vnet[str(lf)] = self._get_vnet(lf)
                if self.config.use_split_xnets:
                    xnet[str(lf)] = {'0': self._get_xnet(lf, first=True), '1': self._get_xnet(lf, first=False)}
                else:
                    xnet[str(lf)] = self._get_xnet(lf, first=True)
        else:
            vnet = self._get_vnet(0)
            if self.config.use_split_xnets:
                xnet = {'0': self._get_xnet(0, first=True), '1': self._get_xnet(0, first=False)}
            else:
                xnet = self._get_xnet(0, first=True)
        return {'xnet': xnet, 'vnet': vnet}

    def init_weights(self, method='xavier_uniform', **kwargs):
        """dynamics.py:333-424 (the methods the configs use)"""
        fn = {'zeros': nn.init.zeros_, 'zero': nn.init.zeros_, 'xavier_uniform': nn.init.xavier_uniform_,
              'xavier_normal': nn.init.xavier_normal_, 'kaiming_normal': nn.init.kaiming_normal_,
              'kaiming_uniform': nn.init.kaiming_uniform_}.get(method)
        if fn is None:
            raise ValueError(f'unknown init method {method!r}')
        with torch.no_grad():
            for m in self.modules():
                if isinstance(m, (nn.Linear, nn.Conv2d)):
                    fn(m.weight)
                    if method in ('zeros', 'zero') and m.bias is not None:
                        nn.init.zeros_(m.bias)

    # ---------------------------------------------------------- save / load
    def save(self, outdir: os.PathLike) -> None:
        netdir = Path(outdir).joinpath('networks')
        netdir.mkdir(exist_ok=True, parents=True)
        self.save_eps(outdir=netdir)
        torch.save(self.state_dict(), netdir.joinpath('dynamics.pt').as_posix())

    def save_eps(self, outdir: os.PathLike) -> None:
        """same file placement as the reference (dynamics.py:544-557, which nests
        a second `networks/` when called from `save`)"""
        netdir = Path(outdir).joinpath('networks')
        netdir.mkdir(exist_ok=True, parents=True)
        xeps = np.array([i.detach().cpu().numpy() for i in self.xeps])
        veps = np.array([i.detach().cpu().numpy() for i in self.veps])
        np.save(netdir.joinpath('xeps.npy'), xeps)
        np.save(netdir.joinpath('veps.npy'), veps)
        np.savetxt(netdir.joinpath('xeps.txt').as_posix(), xeps)
        np.savetxt(netdir.joinpath('veps.txt').as_posix(), veps)

    def load(self, outdir: os.PathLike) -> None:
        netdir = Path(outdir).joinpath('networks')
        self.load_state_dict(torch.load(netdir.joinpath('dynamics.pt'), map_location=self.device))

    def load_eps(self, outdir: os.PathLike) -> dict:
        """dynamics.py:566-582"""
        netdir = Path(outdir).joinpath('networks')
        xe = torch.from_numpy(np.load(netdir.joinpath('xeps.npy')))
        ve = torch.from_numpy(np.load(netdir.joinpath('veps.npy')))
        n = self.config.nleapfrog
        return {'xeps': {str(lf): xe[lf] for lf in range(n)}, 'veps': {str(lf): ve[lf] for lf in range(n)}}

    def restore_eps(self, outdir: os.PathLike) -> None:
        """dynamics.py:584-586 (note the reference's nested `networks/networks` placement, kept by save_eps)"""
        self.assign_eps(self.load_eps(Path(outdir).joinpath('networks')))

    def assign_eps(self, eps) -> None:
        n = self.config.nleapfrog
        if isinstance(eps, dict):
            xe, ve = eps['xeps'], eps['veps']
        elif isinstance(eps, tuple):
            xe = {str(i): eps[0] for i in range(n)}
            ve = {str(i): eps[1] for i in range(n)}
        elif isinstance(eps, float):
            xe = {str(i): eps for i in range(n)}
            ve = {str(i): eps for i in range(n)}
        else:
            raise TypeError
        rg = (not self.config.eps_fixed)
        self.xeps = nn.ParameterList(
            [nn.Parameter(torch.as_tensor(float(xe[str(i)])), requires_grad=rg) for i in range(n)]).to(self.device)
        self.veps = nn.ParameterList(
            [nn.Parameter(torch.as_tensor(float(ve[str(i)])), requires_grad=rg) for i in range(n)]).to(self.device)

    # ------------------------------------------------------------ transitions
    def forward(self, inputs: tuple[Tensor, Tensor]) -> tuple[Tensor, dict]:
        x, beta = inputs
        x = x.to(self._device)
        inputs = (x, beta)
        return self.apply_transition_fb(inputs) if self.config.merge_directions else self.apply_transition(inputs)

    def flatten(self, x: Tensor) -> Tensor:
        return x.reshape(x.shape[0], -1)

    def unflatten(self, x: Tensor) -> Tensor:
        return x.reshape(x.shape[0], *self.xshape[1:])

    def _mix(self, data: dict, sumlogdet_key: bool) -> tuple[Tensor, dict]:
        """accept/reject: out = ma*proposed + mr*init (dynamics.py:632-702) as a
        bit-exact per-chain select in one kernel"""
        ma_, _ = self._get_accept_masks(data['metrics']['acc'])
        init, prop = data['init'], data['proposed']
        xout, vout = ops.accept_mix(ma_, [(init.x, prop.x), (init.v, prop.v)])
        state_out = State(x=xout, v=vout, beta=init.beta)
        mc_states = MonteCarloStates(init=init, proposed=prop, out=state_out)
        if sumlogdet_key:
            data['metrics'].update({'beta': init.beta, 'acc_mask': ma_,
                                    'sumlogdet': ma_ * data['metrics']['sumlogdet'], 'mc_states': mc_states})
        else:
            data['metrics'].update({'acc_mask': ma_, 'mc_states': mc_states})
        return xout, data['metrics']

    def apply_transition_hmc(self, inputs: tuple[Tensor, Tensor], eps: Optional[float] = None,
                             nleapfrog: Optional[int] = None) -> tuple[Tensor, dict]:
        data = self.generate_proposal_hmc(inputs, eps=eps, nleapfrog=nleapfrog)
        return self._mix(data, sumlogdet_key=False)

    def apply_transition_fb(self, inputs: tuple[Tensor, Tensor]) -> tuple[Tensor, dict]:
        data = self.generate_proposal_fb(inputs)
        return self._mix(data, sumlogdet_key=True)

    def apply_transition(self, inputs: tuple[Tensor, Tensor]) -> tuple[Tensor, dict]:
        forward = bool(torch.rand(1) > 0.5)
        data = self.generate_proposal(inputs, forward=forward)
        return self._mix(data, sumlogdet_key=True)

    def apply_transition_both(self, inputs: tuple[Tensor, Tensor]) -> tuple[Tensor, dict]:
        """forward AND backward proposals, a per-chain direction coin, then accept/reject
        (dynamics.py:744-803).  The mixes are per-chain selects done by `k_accept_mix`."""
        x, beta = inputs
        fwd = self.generate_proposal(inputs, forward=True)
        bwd = self.generate_proposal(inputs, forward=False)
        mf_, mb_ = self._get_direction_masks(batch_size=x.shape[0])
        mf_, mb_ = mf_.to(self.device), mb_.to(self.device)
        v_init, xp, vp = ops.accept_mix(mf_, [(bwd['init'].v, fwd['init'].v), (bwd['proposed'].x, fwd['proposed'].x),
                                              (bwd['proposed'].v, fwd['proposed'].v)])
        mfwd, mbwd = fwd['metrics'], bwd['metrics']
        logdetp = mf_ * mfwd['sumlogdet'] + mb_ * mbwd['sumlogdet']
        acc = mf_ * mfwd['acc'] + mb_ * mbwd['acc']
        ma_, mr_ = self._get_accept_masks(acc)
        x_out, v_out = ops.accept_mix(ma_, [(x, xp), (v_init, vp)])
        state_init = State(x=x, v=v_init, beta=beta)
        state_prop = State(x=xp, v=vp, beta=beta)
        state_out = State(x=x_out, v=v_out, beta=beta)
        metrics = {}
        for (key, vf), (_, vb) in zip(mfwd.items(), mbwd.items()):
            if isinstance(vf, Tensor) and vf.dim() >= 1 and vf.shape[-1] == ma_.shape[0]:
                metrics[key] = ma_ * (mf_ * vf + mb_ * vb)
        metrics.update({'acc': acc, 'acc_mask': ma_, 'sumlogdet': ma_ * logdetp,
                        'mc_states': MonteCarloStates(init=state_init, proposed=state_prop, out=state_out)})
        return x_out, metrics

    def random_state(self, beta: float) -> State:
        x = self.g.random(list(self.xshape))
        v = self.g.random_momentum(list(self.xshape))
        return State(x=x, v=v, beta=torch.tensor(beta).to(self.device))

    def test_reversibility(self) -> dict:
        state = self.random_state(beta=1.)
        state_fwd, _ = self.transition_kernel(state, forward=True)
        state_, _ = self.transition_kernel(state_fwd, forward=False)
        dx = (self.flatten(state.x) - self.flatten(state_.x)).abs()
        dv = (self.flatten(state.v) - self.flatten(state_.v)).abs()
        return {'dx': dx.detach().cpu().numpy(), 'dv': dv.detach().cpu().numpy()}

    def _momentum(self, x: Tensor) -> Tensor:
        return self.g.random_momentum([x.shape[0], *self.xshape[1:]])

    def generate_proposal_hmc(self, inputs, eps: Optional[float] = None, nleapfrog: Optional[int] = None) -> dict:
        x, beta = inputs
        x = x.to(self._device)
        init = State(x=x, v=self._momentum(x), beta=beta)
        proposed, metrics = self.transition_kernel_hmc(init, eps=eps, nleapfrog=nleapfrog)
        return {'init': init, 'proposed': proposed, 'metrics': metrics}

    def generate_proposal_fb(self, inputs) -> dict:
        x, beta = inputs
        init = State(x=x, v=self._momentum(x), beta=beta)
        proposed, metrics = self.transition_kernel_fb(State(init.x, init.v, beta))
        return {'init': init, 'proposed': proposed, 'metrics': metrics}

    def generate_proposal(self, inputs, forward: bool) -> dict:
        x, beta = inputs
        init = State(x=x, v=self._momentum(x), beta=beta)
        proposed, metrics = self.transition_kernel(init, forward)
        return {'init': init, 'proposed': proposed, 'metrics': metrics}

    # ------------------------------------------------------------ metrics
    def get_metrics(self, state: State, logdet: Tensor, step: Optional[int] = None,
                    extras: Optional[dict] = None) -> dict:
        energy = self.hamiltonian(state)
        metrics = {'energy': energy, 'logprob': energy - logdet, 'logdet': logdet}
        if extras is not None:
            metrics.update(extras)
        if step is not None:
            metrics.update({'xeps': self.xeps[step], 'veps': self.veps[step]})
        return metrics

    def update_history(self, metrics: dict, history: dict):
        for key, val in metrics.items():
            history.setdefault(key, []).append(val)
        return history

    @staticmethod
    def _stack_history(history: dict) -> dict:
        for key, val in history.items():
            if isinstance(val, list) and isinstance(val[0], Tensor):
                history[key] = torch.stack(val)
        return history

    def _zeros(self, nb: int) -> Tensor:
        return torch.zeros(nb, device=self.device)

    # ------------------------------------------------------------ plain HMC
    def leapfrog_hmc(self, state: State, eps: Optional[float] = None) -> State:
        """one step with two un-merged half kicks (dynamics.py:900-913); used by
        the verbose path -- the fused trajectory kernel is the fast path"""
        eps = self.config.eps if eps is None else eps
        beta = _fbeta(state.beta)
        if self._su3:
            x, v = self.unflatten(state.x), self.unflatten(state.v)
            v1, _ = ops.su3_vupdate(v, self.grad_potential(x, state.beta).detach(), None, None, None, eps, +1)
            xp = ops.su3_update_gauge(x, v1, eps)
            v2, _ = ops.su3_vupdate(v1, self.grad_potential(xp, state.beta).detach(), None, None, None, eps, +1)
            return State(x=xp, v=v2, beta=state.beta)
        x_ = state.x.reshape_as(state.v)
        shape = self.config.latvolume
        v1, _ = ops.u1_vupdate(state.v, ops.u1_force(x_, beta, shape), None, None, None, eps, +1)
        xp = x_ + eps * v1
        v2, _ = ops.u1_vupdate(v1, ops.u1_force(xp, beta, shape), None, None, None, eps, +1)
        return State(x=xp, v=v2, beta=state.beta)

    def transition_kernel_hmc(self, state: State, eps: Optional[float] = None,
                              nleapfrog: Optional[int] = None) -> tuple[State, dict]:
        nb = state.x.shape[0]
        sumlogdet = self._zeros(nb)
        eps = self.config.eps_hmc if eps is None else eps
        nlf = self.config.nleapfrog if not self.config.merge_directions else 2 * self.config.nleapfrog
        if eps is None:
            eps = 1. / nlf
        nleapfrog = nlf if nleapfrog is None else nleapfrog
        beta = _fbeta(state.beta)
        stepwise = self._su3 and getattr(self.lattice, 'c1', 0.0) != 0.0
        # rectangle action: the fused trajectory kernel integrates the plain Wilson force only.
        # nleapfrog = 0: the reference's loop body never runs and the state comes back unchanged (dynamics.py:930-937).
        stepwise = stepwise or nleapfrog <= 0
        if not self._su3:
            # U(1): the whole-trajectory kernel keeps x, v and sin(w) of one chain in shared memory (5 T X elements);
            # lattices beyond 227 KB take the per-step kernels, as every size does in the reference
            T_, X_ = self.config.latvolume
            esize = 8 if state.x.dtype == torch.float64 else 4
            stepwise = stepwise or 5 * T_ * X_ * esize > ops.U1_TRAJECTORY_SMEM_LIMIT
        if stepwise and not self.config.verbose:
            state_ = State(x=state.x, v=state.v, beta=state.beta)
            for _ in range(nleapfrog):
                state_ = self.leapfrog_hmc(state_, eps=eps)
            return state_, {'acc': self.compute_accept_prob(state, state_, sumlogdet), 'sumlogdet': sumlogdet}
        if self.config.verbose:
            state_ = State(x=state.x, v=state.v, beta=state.beta)
            history = self.update_history(self.get_metrics(state_, sumlogdet), {})
            for _ in range(nleapfrog):
                state_ = self.leapfrog_hmc(state_, eps=eps)
                history = self.update_history(self.get_metrics(state_, sumlogdet), history)
            acc = self.compute_accept_prob(state, state_, sumlogdet)
            history.update({'acc': acc, 'sumlogdet': sumlogdet})
            return state_, self._stack_history(history)
        if self._su3:
            xp, vp, en = ops.su3_hmc_trajectory(self.unflatten(state.x), self.unflatten(state.v), beta, eps, nleapfrog)
        else:
            xp, vp, en = ops.u1_hmc_trajectory(state.x, state.v, beta, eps, nleapfrog, shape=self.config.latvolume)
            xp, vp = xp.reshape_as(state.v), vp.reshape_as(state.v)
        prop = State(x=xp, v=vp, beta=state.beta)
        if self._potential_is_wilson():
            dh = (en[:, 0] + en[:, 1]) - (en[:, 2] + en[:, 3]) + sumlogdet
            acc = torch.exp(torch.minimum(dh, torch.zeros_like(dh)))
        else:
            # The reference integrates with the force of its own c1 = 0 lattice (dynamics.py:134,1499) but
            # takes the energies of the accept step from `potential_fn` (dynamics.py:1489-1491): any other
            # potential (improved action, user callable) decides the acceptance, not the kernel's Wilson sums.
            acc = self.compute_accept_prob(state, prop, sumlogdet)
        return prop, {'acc': acc, 'sumlogdet': sumlogdet}

    def _potential_is_wilson(self) -> bool:
        """True when `potential_fn` is the `action` of one of our lattices with the plain Wilson / U(1)
        plaquette action, i.e. exactly the energies the trajectory kernels return (every shipped config)"""
        owner = getattr(self.potential_fn, '__self__', None)
        if getattr(self.potential_fn, '__name__', '') != 'action':
            return False
        if isinstance(owner, LatticeSU3):
            return self._su3 and owner.c1 == 0.0
        return isinstance(owner, LatticeU1) and not self._su3

    # ---------------------------------------------- planar inference sweep (SU(3))
    def _planar_ok(self) -> bool:
        """L2HMC sweep with x, v kept in the kernels' planar layout: inference only (the adjoint
        kernels work on the boundary layout), tensor-core heads, plain Wilson action"""
        if not (self._su3 and self._networks_built) or torch.is_grad_enabled() or self.config.verbose:
            return False
        if getattr(self, 'planar_sweep', 'auto') == 'never' or getattr(self.lattice, 'c1', 0.0) != 0.0:
            return False
        return all(self._fused_heads(self._get_vnet(k)) for k in range(self.config.nleapfrog))

    def _planar_consts(self):
        c = getattr(self, '_planar_cache', None)
        if c is None:
            V = int(np.prod(self.config.latvolume))
            dev = self.masks[0].device
            # planar position (mu, e, site) <- boundary position (mu, site, e)
            perm = torch.arange(4 * V * 9, device=dev).reshape(4, V, 9).permute(0, 2, 1).reshape(-1).contiguous()
            masks = [m.reshape(-1).index_select(0, perm).contiguous() for m in self.masks]
            c = (perm, masks)
            self._planar_cache = c
        return c

    def _eps_tensors(self) -> tuple[Tensor, Tensor]:
        """(x step sizes, v step sizes) = sigmoid(log(.)) of all 2 nlf parameters (dynamics.py:82-83,1270,1394) as ONE
        float64 device vector, so that a sweep costs five tiny element-wise launches for its step sizes instead of
        five per update; entries are handed to the kernels as device pointers (`eps_dev`)"""
        params = list(self.xeps) + list(self.veps)
        t = sigmoid(torch.stack([q.detach().reshape(()) for q in params]).log()).to(torch.float64)
        n = len(self.xeps)
        return t[:n], t[n:]

    def _fused_input(self, vnet, nb: int) -> bool:
        """vnet input layer on the tensor cores (ops.su3_input_layer): dense input, <= 256 units and chains, one of
        the reference's activations; `tensor_core_input`: 'auto' (default) / 'never'"""
        if getattr(self, 'tensor_core_input', 'auto') == 'never' or not vnet.dense_input():
            return False
        V = int(np.prod(self.config.latvolume))
        return (ops.input_layer_supported(4 * V, vnet.units[0], nb) and vnet.input_activation_name() is not None
                and not isinstance(vnet.input_layer.xlayer, nn.modules.lazy.LazyModuleMixin))

    def _transition_kernel_fb_planar(self, state: State) -> tuple[State, dict]:
        """transition_kernel_fb (dynamics.py:956-1029) with the state planar between the two layout conversions at
        its ends; same arithmetic per link and per update as the boundary-layout path.  The schedule exploits what
        the reference recomputes: between two link updates the links do not move, so the force, the vnet inputs and
        -- when the layers share one vnet -- the network outputs of the two momentum updates in between are the
        same: they are evaluated once and both updates run as ONE pass of the heads kernel
        (`l2b_su3_heads_vupdate_pair`, including the v -> -v of the turn-around); the two masked link updates of a
        layer share exp(eps v) and run as one pass too (`l2b_su3_update_gauge_planar_pair`).
        `pair_updates = 'never'` keeps one kernel per update (tests compare the two)."""
        nb = state.x.shape[0]
        perm, pmasks = self._planar_consts()
        beta = _fbeta(state.beta)
        xs = ops.su3_aos_to_soa(self.unflatten(state.x))
        vs = ops.su3_aos_to_soa(self.unflatten(state.v))
        sumlogdet = torch.zeros(nb, dtype=torch.float64, device=xs.device)
        nlf = self.config.nleapfrog
        reuse = self._reuse_force()
        pair = getattr(self, 'pair_updates', 'auto') != 'never' and reuse
        xeps, veps = self._eps_tensors()
        dt = torch.bfloat16 if torch.is_autocast_enabled('cuda') else None
        lm_bufs: list = [None, None]      # link-major activation images, reused across the sweep

        def net_inputs(xs_):
            """force and vnet input producers at the current links"""
            f = ops.su3_force_planar(xs_, beta)
            memo: dict = {'f': f}

            def z_of(vnet):
                if id(vnet) in memo:
                    return memo[id(vnet)]
                ndt = dt or next(vnet.parameters()).dtype
                if ndt == torch.bfloat16 and self._fused_input(vnet, nb):
                    if 'lm' not in memo:
                        lm_bufs[0] = ops.su3_project_vec_planar_lm(xs_, lm_bufs[0])
                        lm_bufs[1] = ops.su3_project_vec_planar_lm(f, lm_bufs[1])
                        memo['lm'] = True
                    z = vnet.hidden_tail(ops.su3_input_layer(lm_bufs[0], lm_bufs[1], vnet.input_pack(), nb))
                else:
                    if ndt not in memo:
                        memo[ndt] = (ops.su3_project_vec_planar(xs_, ndt), ops.su3_project_vec_planar(f, ndt))
                    z = vnet.hidden(memo[ndt])
                memo[id(vnet)] = z
                return z
            return f, z_of

        # the sweep as a flat list: ('v', step, sign) | ('x', step, first_complement, sign) | ('neg',)
        seq: list = []
        for step in range(nlf):                     # _forward_lf
            seq += [('v', step, +1), ('x', step, False, +1), ('v', step, +1)]
        seq.append(('neg',))
        for step in range(nlf):                     # _backward_lf
            r = nlf - step - 1
            seq += [('v', r, -1), ('x', r, True, -1), ('v', r, -1)]
        k = 0
        while k < len(seq):
            op = seq[k]
            if op[0] == 'x':
                _, step, first_c, sign = op
                if pair:
                    xs = ops.su3_update_gauge_planar_pair(xs, vs.reshape(xs.shape), xeps[step], pmasks[step], first_c,
                                                          eps_mult=float(sign))
                else:
                    for comp in (first_c, not first_c):
                        xs = ops.su3_update_gauge_planar(xs, vs.reshape(xs.shape), xeps[step], pmasks[step], comp,
                                                         eps_mult=float(sign))
                k += 1
                continue
            # a run of momentum updates (with at most the turn-around negation inside) at fixed links
            run = []
            while k < len(seq) and seq[k][0] != 'x':
                run.append(seq[k])
                k += 1
            f, z_of = net_inputs(xs)
            i = 0
            while i < len(run):
                if run[i][0] == 'neg':
                    vs = -vs
                    i += 1
                    continue
                _, s1, g1 = run[i]
                vnet1 = self._get_vnet(s1)
                nxt = i + 1
                neg = nxt < len(run) and run[nxt][0] == 'neg'
                if neg:
                    nxt += 1
                second = run[nxt] if nxt < len(run) and run[nxt][0] == 'v' else None
                if pair and second is not None and self._get_vnet(second[1]) is vnet1:
                    vs, ld = ops.su3_heads_vupdate_pair(z_of(vnet1), vnet1.heads_pack(perm), vs.reshape(nb, -1),
                                                        f.reshape(nb, -1), veps[s1], g1, veps[second[1]], second[2],
                                                        negate_between=neg)
                    i = nxt + 1
                else:
                    vs, ld = ops.su3_heads_vupdate(z_of(vnet1), vnet1.heads_pack(perm), vs.reshape(nb, -1),
                                                   f.reshape(nb, -1), veps[s1], g1)
                    i += 1
                sumlogdet = sumlogdet + ld
                if not reuse and i < len(run):
                    f, z_of = net_inputs(xs)          # the reference's schedule: recompute for every update
        xo = ops.su3_soa_to_aos(xs)
        vo = ops.su3_soa_to_aos(vs.reshape(xs.shape))
        out = State(x=xo, v=vo, beta=state.beta)
        acc = self.compute_accept_prob(state, out, sumlogdet)
        return out, {'acc': acc, 'sumlogdet': sumlogdet}

    def transition_kernel_fb(self, state: State) -> tuple[State, dict]:
        if self._planar_ok():
            return self._transition_kernel_fb_planar(state)
        with self._eps_sweep():
            return self._transition_kernel_fb_eager(state)

    def _transition_kernel_fb_eager(self, state: State) -> tuple[State, dict]:
        nb = state.x.shape[0]
        sumlogdet = self._zeros(nb)
        sldf, sldb = torch.zeros_like(sumlogdet), torch.zeros_like(sumlogdet)
        state_ = State(x=state.x, v=state.v, beta=state.beta)
        history: dict = {}
        verbose = self.config.verbose
        if verbose:
            extras = {'sldf': sldf, 'sldb': sldb, 'sld': sumlogdet}
            history = self.update_history(self.get_metrics(state_, sumlogdet, step=0, extras=extras), history)
        for step in range(self.config.nleapfrog):
            state_, logdet = self._forward_lf(step, state_)
            sumlogdet = sumlogdet + logdet
            if verbose:
                sldf = sldf + logdet
                extras = {'sldf': sldf, 'sldb': sldb, 'sld': sumlogdet}
                history = self.update_history(self.get_metrics(state_, sumlogdet, step=step, extras=extras), history)
        state_ = State(state_.x, -state_.v, state_.beta)
        for step in range(self.config.nleapfrog):
            state_, logdet = self._backward_lf(step, state_)
            sumlogdet = sumlogdet + logdet
            if verbose:
                sldb = sldb + logdet
                extras = {'sldf': torch.zeros_like(sldb), 'sldb': sldb, 'sld': sumlogdet}
                history = self.update_history(
                    self.get_metrics(state_, sumlogdet, step=(self.config.nleapfrog - step - 1), extras=extras),
                    history)
        acc = self.compute_accept_prob(state, state_, sumlogdet)
        history.update({'acc': acc, 'sumlogdet': sumlogdet})
        return state_, (self._stack_history(history) if verbose else history)

    def transition_kernel(self, state: State, forward: bool) -> tuple[State, dict]:
        with self._eps_sweep():
            return self._transition_kernel_eager(state, forward)

    def _transition_kernel_eager(self, state: State, forward: bool) -> tuple[State, dict]:
        lf_fn = self._forward_lf if forward else self._backward_lf
        sinit = State(x=state.x, v=state.v, beta=state.beta)
        sumlogdet = self._zeros(state.x.shape[0])
        history: dict = {}
        if self.config.verbose:
            history = self.update_history(self.get_metrics(state, sumlogdet), history)
        for step in range(self.config.nleapfrog):
            state, logdet = lf_fn(step, state)
            sumlogdet = sumlogdet + logdet
            if self.config.verbose:
                history = self.update_history(self.get_metrics(state, sumlogdet, step=step), history)
        # NB: the reference passes the states swapped here (dynamics.py:1053-1057)
        acc = self.compute_accept_prob(state_init=state, state_prop=sinit, sumlogdet=sumlogdet)
        history.update({'acc': acc, 'sumlogdet': sumlogdet})
        return state, (self._stack_history(history) if self.config.verbose else history)

    def compute_accept_prob(self, state_init: State, state_prop: State, sumlogdet: Tensor) -> Tensor:
        """exp(min(H0 - H1 + sumlogdet, 0))   (dynamics.py:1065-1079)"""
        self._fcache = None          # end of a sweep: drop the force / vnet-input cache (`_force`)
        dh = self.hamiltonian(state_init) - self.hamiltonian(state_prop) + sumlogdet
        return torch.exp(torch.minimum(dh, torch.zeros_like(dh)))

    @staticmethod
    def _get_accept_masks(px: Tensor) -> tuple[Tensor, Tensor]:
        acc = (px > torch.rand_like(px)).to(torch.float)
        return acc, torch.ones_like(acc) - acc

    @staticmethod
    def _get_direction_masks(batch_size: int) -> tuple[Tensor, Tensor]:
        fwd = (torch.rand(batch_size) > 0.5).to(torch.float)
        return fwd, torch.ones_like(fwd) - fwd

    def _get_mask(self, step: int) -> tuple[Tensor, Tensor]:
        m = self.masks[step]
        return m, torch.ones_like(m) - m

    def _get_vnet(self, step: int):
        if not self._networks_built:
            return self.vnet
        if self.config.use_separate_networks:
            return self.vnet.get_submodule(str(step))
        return self.vnet

    def _get_xnet(self, step: int, first: bool = False):
        if not self._networks_built:
            return self.xnet
        if self.config.use_separate_networks:
            xnet = self.xnet.get_submodule(str(step))
            if self.config.use_split_xnets:
                return xnet.get_submodule('first') if first else xnet.get_submodule('second')
            return xnet
        return self.xnet

    def group_to_vec(self, x: Tensor, dtype: Optional[torch.dtype] = None) -> Tensor:
        if self._su3 and dtype is not None:
            return self.g.group_to_vec(self.unflatten(x), dtype)
        return self.g.group_to_vec(self.unflatten(x))

    def vec_to_group(self, x: Tensor) -> Tensor:
        x = self.unflatten(x)
        if self._su3:
            return self.g.vec_to_group(x)
        return torch.complex(x[..., 0], x[..., 1])

    def _call_vnet(self, step: int, inputs: tuple[Tensor, Tensor]):
        """(x, force) -> (s, t, q); SU(3) inputs are su3_to_vec(projectSU(.)) of
        BOTH x and the force (dynamics.py:1142-1159)"""
        x, force = inputs
        if self._su3:
            if not self._networks_built:
                return None, None, None       # dummy network: zeros -> plain kick
            vnet = self._get_vnet(step)
            dt = next(vnet.parameters()).dtype      # nets live in torch's default dtype
            if torch.is_autocast_enabled('cuda'):   # cfg 5: bf16 nets; the kernel writes bf16 directly
                dt = torch.get_autocast_dtype('cuda')
            return vnet(self._vnet_vecs(State(x, None, None), force, dt))
        vnet = self._get_vnet(step)
        return vnet((x, force))

    def _call_xnet(self, step: int, inputs: tuple[Tensor, Tensor], first: bool = False):
        x, v = inputs
        xnet = self._get_xnet(step, first)
        if not self._su3:
            x = self.g.group_to_vec(x)
        else:
            x = torch.stack([x.real, x.imag], 1)
            v = torch.stack([v.real, v.imag], 1)
        return xnet((x, v))

    def _stack_as_xy(self, x: Tensor) -> Tensor:
        """[cos(x), sin(x)] on a new last axis (dynamics.py:1137-1140)"""
        return torch.stack([x.cos(), x.sin()], dim=-1).to(self.device)

    @staticmethod
    def complexify(x: Tensor, dim: int = 1) -> Tensor:
        """dynamics.py:1501-1535"""
        assert len(x.shape) >= 2 and x.shape[dim] == 2
        if dim != 1:
            xr, xi = x.transpose(0, dim)
            return torch.complex(xr.transpose(0, dim - 1), xi.transpose(0, dim - 1))
        return torch.complex(x[:, 0], x[:, 1])

    # plain-HMC half updates with the TRAINABLE step sizes (dynamics.py:1244-1264)
    def _update_v_fwd_hmc(self, step: int, state: State) -> Tensor:
        return self._update_v(step, state, +1, hmc=True)[0].v

    def _update_v_bwd_hmc(self, step: int, state: State) -> Tensor:
        return self._update_v(step, state, -1, hmc=True)[0].v

    def _update_x_fwd_hmc(self, step: int, state: State) -> Tensor:
        return self._update_x_hmc(step, state, +1)

    def _update_x_bwd_hmc(self, step: int, state: State) -> Tensor:
        return self._update_x_hmc(step, state, -1)

    def _update_x_hmc(self, step: int, state: State, sign: int) -> Tensor:
        eps_t = self._eps_t(self.xeps[step])
        if self._su3:
            return ag.SU3UpdateGauge.apply(self.unflatten(state.x), self.unflatten(state.v), eps_t.to(torch.float64),
                                           None, sign, None)
        return self.g.update_gauge(state.x.reshape_as(state.v), (sign * eps_t) * state.v)

    def _eps(self, p: Tensor) -> float:
        """sigmoid(log(eps)) == eps / (1 + eps)   (dynamics.py:82-83,1270,1394) as a host
        float (logging / callers that want a number).  The update kernels do NOT use this: they
        read the 0-dim device tensor `_eps_t(p)` directly (include/l2b.h, `eps_dev`).  All 2*nlf
        step sizes are read back with ONE device->host copy and cached until a parameter changes."""
        params = list(self.xeps) + list(self.veps)
        key = tuple((id(q), q._version) for q in params) + (weights_generation(),)
        cache = getattr(self, '_eps_cache', None)
        if cache is None or cache[0] != key:
            vals = sigmoid(torch.stack([q.detach().reshape(()) for q in params]).log()).tolist()
            cache = (key, {id(q): v for q, v in zip(params, vals)})
            self._eps_cache = cache
        if id(p) in cache[1]:
            return cache[1][id(p)]
        return float(sigmoid(p.detach().log()))

    def _eps_t(self, p: Tensor, f64: bool = False) -> Tensor:
        """same, as a 0-dim tensor attached to the graph (trainable step sizes), optionally as float64.  Inside a
        sweep (`_eps_sweep`) all 2 nlf of them come from ONE batched autograd node (`ag.EpsPrime`)."""
        cache = getattr(self, '_eps_cache', None)
        if cache is not None:
            if cache.get('vals') is None:
                params = list(self.xeps) + list(self.veps)
                cache['index'] = {id(q): i for i, q in enumerate(params)}
                cache['vals'] = ag.EpsPrime.apply(*params)
            i = cache['index'].get(id(p))
            if i is not None:
                return cache['vals'][i + len(cache['index']) if f64 else i]
        e = sigmoid(p.log())
        return e.to(torch.float64) if f64 else e

    def _eps_sweep(self):
        """context manager around one sweep of leapfrog layers: the step sizes are transformed once"""
        dyn = self

        class _Ctx:
            def __enter__(self_):
                self_.outer = getattr(dyn, '_eps_cache', None)
                if self_.outer is None:
                    dyn._eps_cache = {}

            def __exit__(self_, *exc):
                if self_.outer is None:
                    dyn._eps_cache = None
                return False
        return _Ctx()

    def _forward_lf(self, step: int, state: State) -> tuple[State, Tensor]:
        m, mb = self._get_mask(step)
        state, logdet = self._update_v_fwd(step, state)
        sumlogdet = logdet
        state, logdet = self._update_x_fwd(step, state, m, first=True)
        sumlogdet = sumlogdet + logdet
        state, logdet = self._update_x_fwd(step, state, mb, first=False)
        sumlogdet = sumlogdet + logdet
        state, logdet = self._update_v_fwd(step, state)
        return state, sumlogdet + logdet

    def _backward_lf(self, step: int, state: State) -> tuple[State, Tensor]:
        step_r = self.config.nleapfrog - step - 1
        m, mb = self._get_mask(step_r)
        state, logdet = self._update_v_bwd(step_r, state)
        sumlogdet = logdet
        state, logdet = self._update_x_bwd(step_r, state, mb, first=False)
        sumlogdet = sumlogdet + logdet
        state, logdet = self._update_x_bwd(step_r, state, m, first=True)
        sumlogdet = sumlogdet + logdet
        state, logdet = self._update_v_bwd(step_r, state)
        return state, sumlogdet + logdet

    def _reuse_force(self) -> bool:
        """Two consecutive v-updates with no x-update in between -- the second half of leapfrog layer i
        and the first half of layer i+1, and the two around the turn-around of the forward/backward
        sweep -- see the same links, so the force and the vnet inputs su3_to_vec(projectSU(.)) of x and
        of the force are the same tensors.  The reference recomputes them (dynamics.py:1187-1228,
        1266-1297); with `reuse_force = 'always'` they are computed once per distinct x (2 nlf + 1
        instead of 4 nlf force evaluations per fb sweep; forward results bit-identical, under autograd
        the shared tensors simply receive the sum of both cotangents).  Default 'always' since round 2
        (tests/test_gpu_reuse_force.py: bit-identical sweeps, force count, gradients); 'never' = recompute
        as the reference does."""
        return getattr(self, 'reuse_force', 'always') == 'always'

    def _force(self, state: State) -> Tensor:
        if not self._reuse_force():
            return self.grad_potential(state.x, state.beta)
        c = getattr(self, '_fcache', None)
        if (c is not None and c['x'] is state.x and c['ver'] == state.x._version and c['beta'] is state.beta
                and c['grad'] == torch.is_grad_enabled()):
            return c['force']
        force = self.grad_potential(state.x, state.beta)
        self._fcache = {'x': state.x, 'ver': state.x._version, 'beta': state.beta, 'grad': torch.is_grad_enabled(),
                        'force': force, 'vecs': {}}
        return force

    def _vnet_vecs(self, state: State, force: Tensor, dt) -> tuple[Tensor, Tensor]:
        """(group_to_vec(x), group_to_vec(force)) in the vnet's dtype (dynamics.py:1154-1156)"""
        c = getattr(self, '_fcache', None) if self._reuse_force() else None
        if c is None or c['force'] is not force:
            return self.group_to_vec(state.x, dt), self.group_to_vec(force, dt)
        if dt not in c['vecs']:
            c['vecs'][dt] = (self.group_to_vec(state.x, dt), self.group_to_vec(force, dt))
        return c['vecs'][dt]

    def _fused_heads(self, vnet) -> bool:
        """SU(3) v-update with the vnet heads on the tensor cores (bf16 tcgen05, fp32
        accumulate) fused with the update itself.  `tensor_core_heads`: 'auto' (default) =
        whenever the nets already run in bf16 (autocast, BASELINE cfg 5), 'always', 'never'."""
        mode = getattr(self, 'tensor_core_heads', 'auto')
        if mode == 'never' or not (self._su3 and self._networks_built):
            return False
        if not ops.heads_supported(vnet.units[-1]):
            return False
        if mode == 'always':
            return True
        return torch.is_autocast_enabled('cuda') and torch.get_autocast_dtype('cuda') == torch.bfloat16

    def _u1_fused(self, net, field: Tensor) -> bool:
        """U(1) inference: the three output heads fused with the update (`l2b_u1_heads_update`);
        `fused_u1_heads` = 'auto' (default: whenever no gradient is being recorded) | 'never'"""
        if not self._networks_built or torch.is_grad_enabled() or torch.is_autocast_enabled('cuda'):
            return False
        if getattr(self, 'fused_u1_heads', 'auto') == 'never':
            return False
        return ops.u1_heads_supported(net.units[-1]) and net.transl.weight.dtype == field.dtype

    def _u1_hidden(self, net, mode: int, x: Tensor, v: Tensor, m: Optional[Tensor]) -> Tensor:
        """hidden vector z of a U(1) net; dense nets: both input Linears (with the cos / sin of the masked links
        for the xnet) in one pass over x and v (`l2b_u1_input_layer`), else the module's own input layer"""
        il = net.input_layer
        if net.dense_input() and ops.u1_input_supported(net.units[0]) and il.xlayer.weight.dtype == v.dtype:
            pre = ops.u1_input_layer(mode, x, v, il.xlayer.weight, il.xlayer.bias, il.vlayer.weight, il.vlayer.bias,
                                     mask=m)
            return net.hidden_from_pre(pre)
        if mode == 1:
            return net.hidden((self.g.group_to_vec(self.unflatten(m) * self.unflatten(x)), v))
        return net.hidden((x, v))

    def _update_v(self, step: int, state: State, sign: int, hmc: bool = False) -> tuple[State, Tensor]:
        """dynamics.py:1266-1297: force, vnet, then the fused epilogue kernel (hmc=True: no
        networks, the plain half kick v -+ eps/2 F of dynamics.py:1244-1254)"""
        force = self._force(state)
        eps = None      # the kernels read the step size from the device tensor (no host round trip)
        if hmc:
            if self._su3:
                return State(state.x, ag.SU3VUpdate.apply(self.unflatten(state.v), self.unflatten(force), None, None,
                                                          None, self._eps_t(self.veps[step], f64=True), sign,
                                                          eps)[0], state.beta), self._zeros(state.x.shape[0])
            f = force.reshape_as(state.v)
            return State(state.x, ag.U1VUpdate.apply(state.v, f, None, None, None, self._eps_t(self.veps[step]), sign,
                                                     eps)[0], state.beta), self._zeros(state.x.shape[0])
        if self._su3 and self._networks_built and self._fused_heads(self._get_vnet(step)):
            vnet = self._get_vnet(step)
            dt = torch.bfloat16 if torch.is_autocast_enabled('cuda') else next(vnet.parameters()).dtype
            z = vnet.hidden(self._vnet_vecs(state, force, dt))
            v, logdet = ag.SU3HeadsVUpdate.apply(z, self.unflatten(state.v), self.unflatten(force),
                                                 self._eps_t(self.veps[step], f64=True), sign, eps, vnet,
                                                 *vnet.head_params())
            return State(state.x, v, state.beta), logdet
        if not self._su3 and self._u1_fused(self._get_vnet(step), state.v):
            vnet = self._get_vnet(step)          # heads + update in one kernel: s, t, q never reach HBM
            z = self._u1_hidden(vnet, 0, state.x, force, None)
            v, logdet = ops.u1_heads_update(0, z, vnet.head_params(), (vnet.nw.s, vnet.nw.t, vnet.nw.q), state.v,
                                            force, self._eps_t(self.veps[step]), sign)
            return State(state.x, v, state.beta), logdet
        s, t, q = self._call_vnet(step, (state.x, force))
        if self._su3:
            v, logdet = ag.SU3VUpdate.apply(self.unflatten(state.v), self.unflatten(force), s, t, q,
                                            self._eps_t(self.veps[step], f64=True), sign, eps)
        else:
            v, logdet = ag.U1VUpdate.apply(state.v, force, s, t, q, self._eps_t(self.veps[step]), sign, eps)
        return State(state.x, v, state.beta), logdet

    def _update_v_fwd(self, step: int, state: State) -> tuple[State, Tensor]:
        return self._update_v(step, state, +1)

    def _update_v_bwd(self, step: int, state: State) -> tuple[State, Tensor]:
        return self._update_v(step, state, -1)

    def _update_x(self, step: int, state: State, m: Tensor, first: bool, sign: int) -> tuple[State, Tensor]:
        """dynamics.py:1386-1477"""
        eps = None      # device-resident step size, as in _update_v
        x = self.unflatten(state.x)
        if self._su3:
            # x' = m*x + exp(+-eps v) @ ((1-m)*x); xnet is never called, logdet = 0
            xn = ag.SU3UpdateGauge.apply(x, self.unflatten(state.v), self._eps_t(self.xeps[step], f64=True),
                                         m, sign, eps)
            return State(x=xn, v=state.v, beta=state.beta), self._zeros(x.shape[0])
        xnet = self._get_xnet(step, first)
        if self._u1_fused(xnet, state.v):
            z = self._u1_hidden(xnet, 1, x, state.v, m)
            xn, logdet = ops.u1_heads_update(1, z, xnet.head_params(), (xnet.nw.s, xnet.nw.t, xnet.nw.q), x, state.v,
                                             self._eps_t(self.xeps[step]), sign, mask=m,
                                             use_ncp=bool(self.config.use_ncp))
            return State(x=xn.reshape(x.shape), v=state.v, beta=state.beta), logdet
        xm_init = self.unflatten(m) * x
        s, t, q = self._call_xnet(step, (xm_init, state.v), first=first)
        xn, logdet = ag.U1XUpdate.apply(x, state.v, s, t, q, m, self._eps_t(self.xeps[step]), sign,
                                        bool(self.config.use_ncp), eps)
        return State(x=xn, v=state.v, beta=state.beta), logdet

    def _update_x_fwd(self, step: int, state: State, m: Tensor, first: bool) -> tuple[State, Tensor]:
        return self._update_x(step, state, m, first, +1)

    def _update_x_bwd(self, step: int, state: State, m: Tensor, first: bool) -> tuple[State, Tensor]:
        return self._update_x(step, state, m, first, -1)

    # ------------------------------------------------------------ energies
    def hamiltonian(self, state: State) -> Tensor:
        return self.kinetic_energy(state.v) + self.potential_energy(state.x, state.beta)

    def kinetic_energy(self, v: Tensor) -> Tensor:
        return self.g.kinetic_energy(self.unflatten(v) if self._su3 else v)

    def potential_energy(self, x: Tensor, beta: Tensor):
        return self.potential_fn(x, beta)

    def grad_potential(self, x: Tensor, beta: Tensor) -> Tensor:
        return self.lattice.grad_action(x, beta)

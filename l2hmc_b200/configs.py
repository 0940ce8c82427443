"""Boundary dataclasses of the hot path, field-compatible with the reference's
`l2hmc/configs.py` (`State` :142-146, `Charges`/`LatticeMetrics` :184-202,
`NetWeight(s)` :278-317, `ConvolutionConfig` :393-434, `NetworkConfig` :437-455,
`DynamicsConfig` :458-520, `LossConfig` :523-538, `InputSpec` :541-571), so that
the Hydra YAML nodes whose `_target_` is `l2hmc.configs.<X>` instantiate these
classes unchanged (`from_target_dict`).  Nothing here touches hydra, mpi4py or
the file system (the reference `mkdir`s at import, configs.py:42-46).
"""
from __future__ import annotations

from dataclasses import asdict, dataclass, field
from typing import Any, Dict, List, Optional, Sequence

import numpy as np


@dataclass
class State:
    x: Any
    v: Any
    beta: Any


@dataclass
class Charges:
    intQ: Any
    sinQ: Any


@dataclass
class LatticeMetrics:
    plaqs: Any
    charges: Charges
    p4x4: Any

    def asdict(self) -> dict:
        return {'plaqs': self.plaqs, 'sinQ': self.charges.sinQ,
                'intQ': self.charges.intQ, 'p4x4': self.p4x4}


class BaseConfig:
    def get_config(self) -> dict:
        return asdict(self)

    def asdict(self) -> dict:
        return asdict(self)

    def __getitem__(self, key):
        return getattr(self, key)

    def to_json(self) -> str:
        """configs.py:157-158"""
        import json
        return json.dumps(self.__dict__)

    def to_dict(self) -> dict:
        """configs.py:166-167"""
        from copy import deepcopy
        return deepcopy(self.__dict__)


def list_to_str(x: list) -> str:
    """dash-joined list, floats with one decimal (configs.py:133-139)"""
    if isinstance(x[0], int):
        return '-'.join(str(int(i)) for i in x)
    if isinstance(x[0], float):
        return '-'.join(f'{i:2.1f}' for i in x)
    return '-'.join(str(i) for i in x)


@dataclass
class NetWeight(BaseConfig):
    """s scales the scaling fn, t the translation, q the transformation
    (configs.py:278-295)."""
    s: float = 1.
    t: float = 1.
    q: float = 1.

    def to_dict(self):
        return {'s': self.s, 't': self.t, 'q': self.q}

    def to_str(self):
        return f's{self.s:2.1f}t{self.t:2.1f}q{self.t:2.1f}'


@dataclass
class NetWeights(BaseConfig):
    x: NetWeight = field(default_factory=lambda: NetWeight(1., 1., 1.))
    v: NetWeight = field(default_factory=lambda: NetWeight(1., 1., 1.))

    def __post_init__(self):
        if not isinstance(self.x, NetWeight):
            self.x = NetWeight(**self.x)
        if not isinstance(self.v, NetWeight):
            self.v = NetWeight(**self.v)

    def to_dict(self):
        return {'x': self.x.to_dict(), 'v': self.v.to_dict()}

    def to_str(self):
        return f'nwx-{self.x.to_str()}-nwv-{self.v.to_str()}'


@dataclass
class ConvolutionConfig(BaseConfig):
    filters: Optional[Sequence[int]] = None
    sizes: Optional[Sequence[int]] = None
    pool: Optional[Sequence[int]] = None

    def __post_init__(self):
        if self.filters is None:
            return
        if self.sizes is None:
            self.sizes = list(len(self.filters) * [2])
        if self.pool is None:
            self.pool = len(self.filters) * [2]
        assert len(self.filters) == len(self.sizes)
        assert len(self.filters) == len(self.pool)

    def to_str(self) -> str:
        """directory-name fragment `conv-<filters>_<sizes>_<pool>` (configs.py:416-434)"""
        if self.filters is None:
            return 'conv-None'
        if len(self.filters) == 0:
            return ''
        parts = [list_to_str(list(self.filters))]
        if self.sizes is not None:
            parts.append(list_to_str(list(self.sizes)))
        if self.pool is not None:
            parts.append(list_to_str(list(self.pool)))
        return 'conv-' + '_'.join(parts)


@dataclass
class NetworkConfig(BaseConfig):
    units: Sequence[int]
    activation_fn: str
    dropout_prob: float
    use_batch_norm: bool = True

    def to_str(self) -> str:
        """`net-<units>_dp-<p>_bn-<flag>` (configs.py:444-448)"""
        units = '-'.join(str(int(u)) for u in self.units)
        return f'net-{units}_dp-{self.dropout_prob:2.1f}_bn-{self.use_batch_norm}'


@dataclass
class DynamicsConfig(BaseConfig):
    nchains: int
    group: str
    latvolume: List[int]
    nleapfrog: int
    eps: float = 0.01
    eps_hmc: float = 0.01
    use_ncp: bool = True
    verbose: bool = True
    eps_fixed: bool = False
    use_split_xnets: bool = True
    use_separate_networks: bool = True
    merge_directions: bool = True

    def to_str(self) -> str:
        latstr = '-'.join([str(i) for i in self.xshape[1:]])
        return '/'.join([self.group, latstr, f'nlf-{self.nleapfrog}',
                         f'xsplit-{self.use_split_xnets}',
                         f'sepnets-{self.use_separate_networks}',
                         f'merge-{self.merge_directions}'])

    def __post_init__(self):
        assert self.group.upper() in ['U1', 'SU3']
        self.latvolume = [int(i) for i in self.latvolume]
        if self.eps_hmc is None:
            self.eps_hmc = 1.0 / self.nleapfrog
        if self.group.upper() == 'U1':
            self.dim = 2
            assert len(self.latvolume) == 2
            self.nt, self.nx = self.latvolume
            self.xshape = (self.nchains, self.dim, *self.latvolume)
            self.vshape = (self.nchains, self.dim, *self.latvolume)
        else:
            self.dim = 4
            assert len(self.latvolume) == 4
            self.link_shape = (3, 3)
            self.vec_shape = 8
            self.nt, self.nx, self.ny, self.nz = self.latvolume
            self.xshape = (self.nchains, self.dim, *self.latvolume, *self.link_shape)
            self.vshape = (self.nchains, self.dim, *self.latvolume, self.vec_shape)
        self.xdim = int(np.prod(self.xshape[1:]))


@dataclass
class LossConfig(BaseConfig):
    use_mixed_loss: bool = False
    charge_weight: float = 0.01
    rmse_weight: float = 0.0
    plaq_weight: float = 0.0
    aux_weight: float = 0.0

    def to_str(self) -> str:
        """configs.py:531-538"""
        return '_'.join([f'qw-{self.charge_weight:2.1f}', f'pw-{self.plaq_weight:2.1f}', f'rw-{self.rmse_weight:2.1f}',
                         f'aw-{self.aux_weight:2.1f}', f'mixed-{self.use_mixed_loss}'])


@dataclass
class InputSpec(BaseConfig):
    xshape: Sequence[int]
    xnet: Optional[Dict[str, Any]] = None
    vnet: Optional[Dict[str, Any]] = None

    def to_str(self) -> str:
        """configs.py:547-548"""
        return '-'.join(str(i) for i in self.xshape)

    def __post_init__(self):
        if len(self.xshape) == 2:
            self.xdim = self.xshape[-1]
            self.vshape = self.xshape
            self.vdim = self.xshape[-1]
        elif len(self.xshape) > 2:
            self.xdim = int(np.prod(self.xshape[1:]))
            lat_shape = self.xshape[:-2]
            vd = (self.xshape[-1] ** 2) - 1
            self.vshape = (*lat_shape, vd)
            self.vdim = int(np.prod(self.vshape[1:]))
        else:
            raise ValueError(f'Invalid `xshape`: {self.xshape}')
        if self.xnet is None:
            self.xnet = {'x': self.xshape, 'v': self.xshape}
        if self.vnet is None:
            self.vnet = {'x': self.xshape, 'v': self.xshape}


def get_input_spec(config: DynamicsConfig) -> InputSpec:
    """`BaseTrainer.get_input_spec` (trainers/trainer.py:292-309)."""
    xdim = config.xdim
    xshape = config.xshape
    if config.group.upper() == 'U1':
        return InputSpec(xshape=xshape,
                         xnet={'x': [xdim, 2], 'v': [xdim, ]},
                         vnet={'x': [xdim, ], 'v': [xdim, ]})
    vdim = int(np.prod(config.vshape[1:]))
    return InputSpec(xshape=xshape,
                     xnet={'x': [vdim, ], 'v': [vdim, ]},
                     vnet={'x': [vdim, ], 'v': [vdim, ]})


_TARGETS = {
    'DynamicsConfig': DynamicsConfig, 'NetworkConfig': NetworkConfig,
    'ConvolutionConfig': ConvolutionConfig, 'NetWeights': NetWeights,
    'NetWeight': NetWeight, 'LossConfig': LossConfig, 'InputSpec': InputSpec,
}


def from_target_dict(node: dict):
    """Instantiate a Hydra-style node `{_target_: l2hmc.configs.X, ...}`
    (conf/dynamics/*.yaml, conf/network/*.yaml, ...) without hydra."""
    node = dict(node)
    target = node.pop('_target_').rsplit('.', 1)[-1]
    if target not in _TARGETS:
        raise KeyError(f'{target} is outside the hot-path boundary (SURVEY 8b)')
    for k, v in list(node.items()):
        if isinstance(v, dict) and '_target_' in v:
            node[k] = from_target_dict(v)
    return _TARGETS[target](**node)

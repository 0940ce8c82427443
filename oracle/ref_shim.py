"""oracle/ref_shim.py -- TEST INFRASTRUCTURE, not product code.

Imports the UNMODIFIED reference hot path from `oracle/_ref/src/l2hmc` (made by
`oracle/make_ref.py`) in an environment that lacks mpi4py / enrich / hydra /
omegaconf.  Only import-time dependencies are stubbed; every arithmetic module
(`group`, `lattice`, `dynamics`, `network`, `loss`) is the reference's own.

Usage (default dtype MUST be chosen before the import: the reference freezes
`TORCH_DTYPE`-typed constants at import, `group/su3/pytorch/utils.py:28-36`)::

    from oracle.ref_shim import load_reference
    ref = load_reference(torch.float64)        # -> namespace of classes
    lat = ref.LatticeSU3(2, [4, 4, 4, 4])

Only `tests/`, `oracle/make_golden.py` and `bench.py`'s reference / cpu_baseline
arm may import this module.
"""
from __future__ import annotations

import contextlib
import io
import logging
import sys
import types
from pathlib import Path
from types import SimpleNamespace

HERE = Path(__file__).resolve().parent
REF_SRC = HERE / '_ref' / 'src'


def available() -> bool:
    return (REF_SRC / 'l2hmc' / 'configs.py').exists()


def _mod(name: str, **kw) -> types.ModuleType:
    m = types.ModuleType(name)
    m.__dict__.update(kw)
    sys.modules[name] = m
    return m


class _Comm:  # stands in for mpi4py.MPI.COMM_WORLD (l2hmc/__init__.py:22-23)
    def Get_rank(self):
        return 0

    def Get_size(self):
        return 1

    def bcast(self, x, root=0):
        return x


class _ConfigStore:  # hydra.core.config_store.ConfigStore (configs.py:1146-1150)
    _inst = None

    @classmethod
    def instance(cls):
        cls._inst = cls._inst or cls()
        return cls._inst

    def store(self, **kw):
        pass


def _install_stubs() -> None:
    try:
        import mpi4py  # noqa: F401
    except Exception:
        _mod('mpi4py').MPI = _mod('mpi4py.MPI', COMM_WORLD=_Comm())
    try:
        import enrich.handler  # noqa: F401
    except Exception:
        _mod('enrich')
        _mod('enrich.handler', RichHandler=logging.StreamHandler)
    try:
        import hydra.core.config_store  # noqa: F401
    except Exception:
        _mod('hydra')
        _mod('hydra.core')
        _mod('hydra.core.config_store', ConfigStore=_ConfigStore)
    try:
        import omegaconf  # noqa: F401
    except Exception:
        _mod('omegaconf', DictConfig=dict)


_LOADED: dict = {}


def load_reference(default_dtype=None, quiet: bool = True) -> SimpleNamespace:
    """Import the reference hot path.  The FIRST call fixes the default dtype
    the reference's module constants are frozen with (fp64 runs must pass
    torch.float64 here before anything else imported `l2hmc`)."""
    import torch
    if not available():
        raise RuntimeError(
            'oracle/_ref is missing: run `python oracle/make_ref.py` where '
            '/root/reference is mounted')
    if default_dtype is not None:
        if _LOADED and _LOADED['dtype'] != default_dtype:
            raise RuntimeError(
                'reference already imported with default dtype '
                f'{_LOADED["dtype"]}; use a fresh process for {default_dtype}')
        torch.set_default_dtype(default_dtype)
    if _LOADED:
        return _LOADED['ns']
    _install_stubs()
    if str(REF_SRC) not in sys.path:
        sys.path.insert(0, str(REF_SRC))
    sink = io.StringIO() if quiet else sys.stdout
    with contextlib.redirect_stdout(sink):
        import l2hmc  # noqa: F401  (prints "Using device")
        # configs.py:22 imports l2hmc.utils.dist (mpi4py/horovod/deepspeed
        # plumbing; only `query_environment` is referenced, configs.py:217).
        utils = _mod('l2hmc.utils')
        utils.__path__ = []  # mark as package
        utils.dist = _mod(
            'l2hmc.utils.dist',
            query_environment=lambda: {
                'rank': 0, 'local_rank': 0, 'world_size': 1})
        import l2hmc.configs as cfgs
        from l2hmc.dynamics.pytorch.dynamics import Dynamics, State
        from l2hmc.lattice.su3.pytorch.lattice import LatticeSU3
        from l2hmc.lattice.u1.pytorch.lattice import LatticeU1
        from l2hmc.group.su3.pytorch.group import SU3
        from l2hmc.group.u1.pytorch.group import U1Phase
        import l2hmc.group.su3.pytorch.utils as su3utils
        from l2hmc.network.pytorch.network import NetworkFactory, LeapfrogLayer
        try:
            from l2hmc.loss.pytorch.loss import LatticeLoss
        except Exception:  # loss pulls optional deps in some versions
            LatticeLoss = None
    ns = SimpleNamespace(
        l2hmc=l2hmc, cfgs=cfgs, Dynamics=Dynamics, State=State,
        LatticeSU3=LatticeSU3, LatticeU1=LatticeU1, SU3=SU3, U1Phase=U1Phase,
        su3utils=su3utils, NetworkFactory=NetworkFactory,
        LeapfrogLayer=LeapfrogLayer, LatticeLoss=LatticeLoss,
        DynamicsConfig=cfgs.DynamicsConfig, NetworkConfig=cfgs.NetworkConfig,
        ConvolutionConfig=cfgs.ConvolutionConfig, NetWeights=cfgs.NetWeights,
        NetWeight=cfgs.NetWeight, InputSpec=cfgs.InputSpec,
    )
    _LOADED.update(ns=ns, dtype=torch.get_default_dtype())
    return ns

"""`Lattice` base class: shape bookkeeping shared by LatticeU1 / LatticeSU3
(reference `lattice/lattice.py:21-227`)."""
from __future__ import annotations

from typing import Any, Optional

import numpy as np

from ..configs import Charges
from ..group.group import Group


class Lattice:
    def __init__(self, group: Group, nchains: int, shape: list[int]) -> None:
        self.g = group
        self.link_shape = self.g._shape
        self.xshape = [self.g._dim, *shape]
        if len(self.g._shape) > 1:
            self.xshape.extend(self.g._shape)
        self.dim = self.g._dim
        self._shape = [nchains, *self.xshape]
        self.nchains = nchains
        self._lattice_shape = list(shape)
        self.volume = int(np.prod(shape))

    def draw_batch(self) -> Any:
        return self.g.random(list(self._shape))

    def random(self) -> Any:
        return self.g.random(list(self._shape))

    def random_momentum(self) -> Any:
        return self.g.random_momentum(list(self._shape))

    def update_link(self, x: Any, p: Any) -> Any:
        return self.g.update_gauge(x, p)

    def potential_energy(self, x: Any, beta: Any) -> Any:
        return self.action(x, beta)

    def unnormalized_log_prob(self, x: Any, beta: Any) -> Any:
        return self.action(x=x, beta=beta)

    # subclasses provide wilson_loops/_plaqs/_charges/_sin_charges/_int_charges
    def plaqs(self, x: Optional[Any] = None, wloops: Optional[Any] = None) -> Any:
        if wloops is None:
            assert x is not None
            wloops = self.wilson_loops(x)
        return self._plaqs(wloops)

    def charges(self, x: Optional[Any] = None, wloops: Optional[Any] = None) -> Charges:
        if wloops is None:
            assert x is not None
            wloops = self.wilson_loops(x)
        return self._charges(wloops=wloops)

    def sin_charges(self, x: Optional[Any] = None, wloops: Optional[Any] = None) -> Any:
        if wloops is None:
            assert x is not None
            wloops = self.wilson_loops(x)
        return self._sin_charges(wloops)

    def int_charges(self, x: Optional[Any] = None, wloops: Optional[Any] = None) -> Any:
        if wloops is None:
            assert x is not None
            wloops = self.wilson_loops(x)
        return self._int_charges(wloops)

"""``StackedDistributedArray`` (pylops_mpi/DistributedArray.py:962-1154): a list of ``DistributedArray`` with
element-wise arithmetic, ``dot`` and ``norm`` across the components ("next" row f3).  The stacked operators follow
the reference's file layout: ``MPIStackedLinearOperator`` in StackedLinearOperator.py, ``MPIStackedBlockDiag`` /
``MPIStackedVStack`` in basicoperators/BlockDiag.py / VStack.py, ``MPIGradient`` in basicoperators/Gradient.py."""
from __future__ import annotations

from typing import List, Optional

import numpy as np
import torch

from .comm import COMM_WORLD, resolve
from .DistributedArray import DistributedArray

__all__ = ["StackedDistributedArray"]


class StackedDistributedArray:
    def __init__(self, distarrays: List[DistributedArray], base_comm=COMM_WORLD):
        self.distarrays = list(distarrays)
        self.narrays = len(self.distarrays)
        self.base_comm = resolve(base_comm)
        self.rank, self.size = self.base_comm.Get_rank(), self.base_comm.Get_size()

    def __getitem__(self, index):
        return self.distarrays[index]

    def __setitem__(self, index, value):
        self.distarrays[index][:] = value

    def asarray(self) -> torch.Tensor:
        return torch.cat([d.asarray().reshape(-1) for d in self.distarrays])

    def _check_stacked_size(self, other):
        if self.narrays != other.narrays:
            raise ValueError("Stacked array must be composed the same number of of distributed arrays")
        for a, b in zip(self.distarrays, other.distarrays):
            if a.global_shape != b.global_shape:
                raise ValueError(f"Stacked array have different global shape {a.global_shape} and {b.global_shape}")

    def __neg__(self):
        return StackedDistributedArray([-d for d in self.distarrays], self.base_comm)

    def __add__(self, x):
        return self.add(x)

    def __iadd__(self, x):
        return self.iadd(x)

    def __sub__(self, x):
        return self.__add__(-x)

    def __isub__(self, x):
        return self.__iadd__(-x)

    def __mul__(self, x):
        return self.multiply(x)

    __rmul__ = __mul__

    def add(self, other):
        self._check_stacked_size(other)
        return StackedDistributedArray([a + b for a, b in zip(self.distarrays, other.distarrays)], self.base_comm)

    def iadd(self, other):
        self._check_stacked_size(other)
        for a, b in zip(self.distarrays, other.distarrays):
            a += b
        return self

    def multiply(self, other):
        if isinstance(other, StackedDistributedArray):
            self._check_stacked_size(other)
            return StackedDistributedArray([a * b for a, b in zip(self.distarrays, other.distarrays)], self.base_comm)
        return StackedDistributedArray([a * other for a in self.distarrays], self.base_comm)

    def dot(self, other, vdot: bool = False):
        self._check_stacked_size(other)
        return np.sum([a.dot(b, vdot=vdot) for a, b in zip(self.distarrays, other.distarrays)], axis=0)

    def norm(self, ord: Optional[int] = None):
        norms = np.array([d.norm(ord)[0] for d in self.distarrays])
        ord = 2 if ord is None else ord
        if ord in ("fro", "nuc"):
            raise ValueError(f"norm-{ord} not possible for vectors")
        if ord == 0:
            return np.array([np.sum(norms)])
        if ord == np.inf:
            return np.array([np.max(norms)])
        if ord == -np.inf:
            return np.array([np.min(norms)])
        return np.array([np.power(np.sum(np.power(norms, ord)), 1.0 / ord)])

    def conj(self):
        return StackedDistributedArray([d.conj() for d in self.distarrays], self.base_comm)

    def copy(self):
        return StackedDistributedArray([d.copy() for d in self.distarrays], self.base_comm)

    def empty_like(self):
        return StackedDistributedArray([d.empty_like() for d in self.distarrays], self.base_comm)

    def __repr__(self):
        return f"<StackedDistributedArray with {self.narrays} distributed arrays: \n" + \
            "\n".join(repr(d) for d in self.distarrays)


_MOVED = {"MPIStackedLinearOperator": ".StackedLinearOperator", "MPIStackedBlockDiag": ".basicoperators.BlockDiag",
          "MPIStackedVStack": ".basicoperators.VStack", "MPIGradient": ".basicoperators.Gradient"}


def __getattr__(name):
    # the stacked operators used to live in this module: resolve the old import path lazily (no import cycle)
    if name in _MOVED:
        import importlib
        return getattr(importlib.import_module(_MOVED[name], __package__), name)
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")

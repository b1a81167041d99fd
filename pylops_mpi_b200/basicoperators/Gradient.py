"""``MPIGradient`` (pylops_mpi/basicoperators/Gradient.py:21-119): first derivative along every axis of a model
distributed over axis 0, stacked -- the distributed axis is :class:`MPIFirstDerivative` (peer-halo stencil kernel),
the other axes are rank-local batched stencils (``b2_derivative_axis``) inside an ``MPIBlockDiag``."""
from __future__ import annotations

import numpy as np

from ..comm import COMM_WORLD, resolve
from ..DistributedArray import Partition, local_split
from ..local import FirstDerivative
from .BlockDiag import MPIBlockDiag
from .FirstDerivative import MPIFirstDerivative
from .VStack import MPIStackedVStack

__all__ = ["MPIGradient"]


class MPIGradient(MPIStackedVStack):
    """Gradient.py:21-119: first derivative along every axis; axis 0 is the distributed
    :class:`MPIFirstDerivative`, the other axes are rank-local batched stencils in an MPIBlockDiag."""

    def __init__(self, dims, sampling=1, edge: bool = False, kind: str = "centered", base_comm=COMM_WORLD,
                 dtype="float64"):
        base_comm = resolve(base_comm)
        self.dims = tuple(int(d) for d in (dims if np.ndim(dims) else (dims,)))
        ndims = len(self.dims)
        self.sampling = tuple(sampling) if np.ndim(sampling) else (sampling,) * ndims
        self.edge, self.kind = edge, kind
        local_dims = local_split(tuple(self.dims), base_comm, Partition.SCATTER, axis=0)
        ops = [MPIFirstDerivative(dims=self.dims, sampling=self.sampling[0], kind=kind, edge=edge,
                                  base_comm=base_comm, dtype=np.dtype(dtype))]
        for iax in range(1, ndims):
            ops.append(MPIBlockDiag([FirstDerivative(dims=local_dims, axis=iax, sampling=self.sampling[iax],
                                                     edge=edge, kind=kind, dtype=np.dtype(dtype))],
                                    base_comm=base_comm))
        super().__init__(ops, base_comm=base_comm, dtype=np.dtype(dtype))

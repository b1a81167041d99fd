"""``MPILaplacian`` (pylops_mpi/basicoperators/Laplacian.py:14-126): weighted sum of second
derivatives; the term along axis 0 is the distributed :class:`MPISecondDerivative`, the terms along
the other axes are rank-local batched stencils inside an :class:`MPIBlockDiag`."""
from __future__ import annotations

from typing import Tuple

import numpy as np

from ..comm import COMM_WORLD, resolve
from ..DistributedArray import DistributedArray, Partition, local_split
from ..LinearOperator import MPILinearOperator
from ..local import SecondDerivative
from .BlockDiag import MPIBlockDiag
from .SecondDerivative import MPISecondDerivative


class MPILaplacian(MPILinearOperator):
    def __init__(self, dims, axes=(-2, -1), weights: Tuple[float, ...] = (1, 1),
                 sampling: Tuple[float, ...] = (1, 1), edge: bool = False, kind: str = "centered",
                 base_comm=COMM_WORLD, dtype=np.float64):
        self.dims = tuple(int(d) for d in dims)
        nd = len(self.dims)
        norm_axes = []
        for ax in axes:
            if not -nd <= ax < nd:
                raise np.exceptions.AxisError(ax, nd)
            norm_axes.append(ax % nd)
        if not (len(norm_axes) == len(weights) == len(sampling)):
            raise ValueError("axes, weights, and sampling have different size")
        self.axes = tuple(norm_axes)
        self.weights = weights
        self.sampling = sampling
        self.edge = edge
        self.kind = kind
        self.dtype = np.dtype(dtype)
        self.base_comm = resolve(base_comm)
        l2op = self._calc_l2op()
        super().__init__(shape=l2op.shape, dtype=self.dtype, base_comm=self.base_comm)
        self.Op_ = l2op           # (MPILinearOperator.Op is reserved for rank-local wrapped operators)

    def _matvec(self, x: DistributedArray) -> DistributedArray:
        return self.Op_ @ x

    def _rmatvec(self, x: DistributedArray) -> DistributedArray:
        return self.Op_.H @ x

    def _term(self, ax, samp):
        if ax == 0:
            return MPISecondDerivative(dims=self.dims, sampling=samp, kind=self.kind, edge=self.edge,
                                       base_comm=self.base_comm, dtype=self.dtype)
        local_dims = local_split(tuple(self.dims), self.base_comm, Partition.SCATTER, axis=0)
        return MPIBlockDiag(ops=[SecondDerivative(dims=local_dims, axis=ax, sampling=samp, kind=self.kind,
                                                  edge=self.edge, dtype=self.dtype)], base_comm=self.base_comm)

    def _calc_l2op(self):
        l2op = self.weights[0] * self._term(self.axes[0], self.sampling[0])
        for ax, samp, weight in zip(self.axes[1:], self.sampling[1:], self.weights[1:]):
            l2op += weight * self._term(ax, samp)
        return l2op

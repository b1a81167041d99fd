"""``MPIVStack`` (pylops_mpi/basicoperators/VStack.py:21-149): forward applies
each rank's operators to the BROADCAST model (SCATTER data, no comm); adjoint
sums the per-rank contributions with an NCCL Allreduce over NVLink into a
BROADCAST model (VStack.py:146-148)."""
from __future__ import annotations

from typing import Sequence

import numpy as np

from ..comm import COMM_WORLD, resolve, SUM
from ..Distributed import allreduce_
from ..DistributedArray import DistributedArray, Partition
from ..LinearOperator import MPILinearOperator, _get_dtype
from ..local import apply_into
from ..StackedArray import StackedDistributedArray
from ..StackedLinearOperator import MPIStackedLinearOperator
from ..utils.decorators import reshaped


class MPIVStack(MPILinearOperator):
    def __init__(self, ops: Sequence, base_comm=COMM_WORLD, dtype=None):
        base_comm = resolve(base_comm)
        self.ops = ops
        nops = np.zeros(len(self.ops), dtype=np.int64)
        for iop, oper in enumerate(self.ops):
            nops[iop] = oper.shape[0]
        self.nops = int(nops.sum())
        info = base_comm.allgather((self.nops, [int(oper.shape[1]) for oper in self.ops]))
        self.local_shapes_n = [(int(i[0]),) for i in info]
        mops = [m for i in info for m in i[1]]
        if len(set(mops)) > 1:
            raise ValueError("Operators have different number of columns")
        self.mops = int(mops[0])
        self.nnops = np.insert(np.cumsum(nops), 0, 0)
        shape = (sum(s[0] for s in self.local_shapes_n), self.mops)
        dtype = _get_dtype(self.ops) if dtype is None else np.dtype(dtype)
        super().__init__(shape=shape, dtype=dtype, base_comm=base_comm)

    def _matvec(self, x: DistributedArray) -> DistributedArray:
        if x.partition not in [Partition.BROADCAST, Partition.UNSAFE_BROADCAST]:
            raise ValueError(f"x should have partition={Partition.BROADCAST},{Partition.UNSAFE_BROADCAST}"
                             f"Got  {x.partition} instead...")
        y = DistributedArray(global_shape=self.shape[0], base_comm=x.base_comm,
                             local_shapes=self.local_shapes_n, dtype=self.dtype, _trusted=True)
        for iop, oper in enumerate(self.ops):
            oi = y.local_array[self.nnops[iop]:self.nnops[iop + 1]]
            apply_into(oper, x.local_array, oi, False)
        return y

    @reshaped(forward=False, stacking=True)
    def _rmatvec(self, x: DistributedArray) -> DistributedArray:
        y = DistributedArray(global_shape=self.shape[1], base_comm=x.base_comm,
                             partition=Partition.BROADCAST, dtype=self.dtype)
        acc = y.local_array
        tmp = None
        for iop, oper in enumerate(self.ops):
            xi = x.local_array[self.nnops[iop]:self.nnops[iop + 1]]
            if iop == 0:
                apply_into(oper, xi, acc, True)
            else:
                tmp = oper.rmatvec(xi)
                if tmp.dtype != acc.dtype:        # mixed-dtype stack: accumulate in the operator's result dtype
                    if tmp.dtype.is_complex and not acc.dtype.is_complex:
                        tmp = tmp.real            # NumPy __setitem__ cast of the reference (VStack.py:144-145)
                    tmp = tmp.to(acc.dtype)
                y._lincomb(1.0, tmp.reshape(-1).contiguous(), 1.0, acc, out=acc)
        if len(self.ops) == 0:
            acc.zero_()
        allreduce_(x.base_comm, acc, SUM)
        return y


class MPIStackedVStack(MPIStackedLinearOperator):
    """VStack.py:152-201: operators applied one after the other to the same model, outputs stacked."""

    def __init__(self, ops: Sequence[MPILinearOperator], base_comm=COMM_WORLD, dtype=None):
        self.ops = ops
        if len(set(op.shape[1] for op in ops)) > 1:
            raise ValueError("Operators have different number of columns")
        shape = (int(sum(op.shape[0] for op in ops)), ops[0].shape[1])
        dtype = _get_dtype(self.ops) if dtype is None else np.dtype(dtype)
        super().__init__(shape=shape, dtype=dtype, base_comm=base_comm)

    def _matvec(self, x: DistributedArray) -> StackedDistributedArray:
        return StackedDistributedArray([oper.matvec(x) for oper in self.ops], self.base_comm)

    def _rmatvec(self, x: StackedDistributedArray) -> DistributedArray:
        y = self.ops[0].rmatvec(x[0])
        for xx, oper in zip(x[1:], self.ops[1:]):
            y += oper.rmatvec(xx)
        return y

"""``MPIBlockDiag`` (pylops_mpi/basicoperators/BlockDiag.py:16-143): each rank
applies its own list of rank-local operators to its slice of the model; no
communication in the apply."""
from __future__ import annotations

from typing import List, Optional, Sequence

import numpy as np
import torch

from ..comm import COMM_WORLD, resolve
from ..DistributedArray import DistributedArray
from ..LinearOperator import MPILinearOperator, _get_dtype
from ..local import apply_into
from ..StackedArray import StackedDistributedArray
from ..StackedLinearOperator import MPIStackedLinearOperator
from ..utils.decorators import reshaped


def _apply_ops(ops, x: torch.Tensor, bounds, out: torch.Tensor, out_bounds, adjoint: bool):
    for iop, oper in enumerate(ops):
        xi = x[bounds[iop]:bounds[iop + 1]]
        oi = out[out_bounds[iop]:out_bounds[iop + 1]]
        apply_into(oper, xi, oi, adjoint)      # b200 local ops write in place (dtype-checked), others via a temp


class MPIBlockDiag(MPILinearOperator):
    def __init__(self, ops: Sequence, base_comm=COMM_WORLD, mask: Optional[List[int]] = None,
                 dtype=None):
        base_comm = resolve(base_comm)
        self.ops = ops
        self.mask = mask
        mops = np.zeros(len(self.ops), dtype=np.int64)
        nops = np.zeros(len(self.ops), dtype=np.int64)
        for iop, oper in enumerate(self.ops):
            nops[iop] = oper.shape[0]
            mops[iop] = oper.shape[1]
        self.mops = int(mops.sum())
        self.nops = int(nops.sum())
        # one host allgather at construction (BlockDiag.py:112-117 does four collectives)
        both = base_comm.allgather((self.mops, self.nops))
        self.local_shapes_m = [(int(b[0]),) for b in both]
        self.local_shapes_n = [(int(b[1]),) for b in both]
        self.nnops = np.insert(np.cumsum(nops), 0, 0)
        self.mmops = np.insert(np.cumsum(mops), 0, 0)
        shape = (sum(s[0] for s in self.local_shapes_n), sum(s[0] for s in self.local_shapes_m))
        dtype = _get_dtype(ops) if dtype is None else np.dtype(dtype)
        super().__init__(shape=shape, dtype=dtype, base_comm=base_comm)

    @reshaped(forward=True, stacking=True)
    def _matvec(self, x: DistributedArray) -> DistributedArray:
        y = DistributedArray(global_shape=self.shape[0], base_comm=x.base_comm,
                             local_shapes=self.local_shapes_n, mask=self.mask, dtype=self.dtype,
                             _trusted=True)       # shapes validated once at construction
        _apply_ops(self.ops, x.local_array, self.mmops, y.local_array, self.nnops, False)
        return y

    @reshaped(forward=False, stacking=True)
    def _rmatvec(self, x: DistributedArray) -> DistributedArray:
        y = DistributedArray(global_shape=self.shape[1], base_comm=x.base_comm,
                             local_shapes=self.local_shapes_m, mask=self.mask, dtype=self.dtype,
                             _trusted=True)
        _apply_ops(self.ops, x.local_array, self.nnops, y.local_array, self.mmops, True)
        return y


class MPIStackedBlockDiag(MPIStackedLinearOperator):
    """BlockDiag.py:146-204: operator i acts on stacked component i."""

    def __init__(self, ops: Sequence[MPILinearOperator], base_comm=COMM_WORLD, dtype=None):
        self.ops = ops
        dtype = _get_dtype(self.ops) if dtype is None else np.dtype(dtype)
        shape = (int(sum(op.shape[0] for op in ops)), int(sum(op.shape[1] for op in ops)))
        super().__init__(shape=shape, dtype=dtype, base_comm=base_comm)

    def _matvec(self, x: StackedDistributedArray) -> StackedDistributedArray:
        return StackedDistributedArray([oper.matvec(xx) for xx, oper in zip(x.distarrays, self.ops)], self.base_comm)

    def _rmatvec(self, x: StackedDistributedArray) -> StackedDistributedArray:
        return StackedDistributedArray([oper.rmatvec(xx) for xx, oper in zip(x.distarrays, self.ops)], self.base_comm)

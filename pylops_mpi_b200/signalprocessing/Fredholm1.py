"""``MPIFredholm1`` (pylops_mpi/signalprocessing/Fredholm1.py:14-171): batched
(per frequency slice) dense product with the kernel ``G`` split over ranks along
the slice axis; BROADCAST model in, BROADCAST data out via an NCCL Allgather(v)
written straight into the output buffer."""
from __future__ import annotations

import numpy as np
import torch

from .. import _lib
from ..comm import COMM_WORLD, resolve
from ..Distributed import allgatherv
from ..DistributedArray import DistributedArray, Partition
from ..LinearOperator import MPILinearOperator


class MPIFredholm1(MPILinearOperator):
    def __init__(self, G, nz: int = 1, saveGt: bool = False, usematmul: bool = True,
                 base_comm=COMM_WORLD, dtype="float64") -> None:
        base_comm = resolve(base_comm)
        self.nz = int(nz)
        if not isinstance(G, torch.Tensor):
            G = torch.as_tensor(np.asarray(G))
        self.nsl, self.nx, self.ny = (int(s) for s in G.shape)
        self.nsls = base_comm.allgather(self.nsl)
        if base_comm.Get_rank() == 0 and 1 in self.nsls:
            raise NotImplementedError(f'All ranks must have at least 2 or more '
                                      f'elements in the first dimension: '
                                      f'local split is instead {self.nsls}...')
        nslstot = int(sum(self.nsls))
        self.islstart = np.insert(np.cumsum(self.nsls)[:-1], 0, 0)
        self.islend = np.cumsum(self.nsls)
        self.dims = (nslstot, self.ny, self.nz)
        self.dimsd = (nslstot, self.nx, self.nz)
        shape = (int(np.prod(self.dimsd)), int(np.prod(self.dims)))
        super().__init__(shape=shape, dtype=np.dtype(dtype), base_comm=base_comm)
        _lib.ctx()
        self._tdtype = _lib.torch_dtype(dtype)
        self.G = G.to(device="cuda", dtype=self._tdtype).contiguous()
        # saveGt / usematmul change how the reference evaluates the adjoint (:147-167),
        # not its result; the batched kernel applies G^H on the fly.
        self.saveGt = saveGt
        self.usematmul = usematmul

    def _apply(self, x: DistributedArray, adjoint: bool) -> DistributedArray:
        if x.partition not in [Partition.BROADCAST, Partition.UNSAFE_BROADCAST]:
            raise ValueError(f"x should have partition={Partition.BROADCAST},{Partition.UNSAFE_BROADCAST}"
                             f"Got  {x.partition} instead...")
        rank = self.rank
        nin, nout = (self.nx, self.ny) if adjoint else (self.ny, self.nx)
        y = DistributedArray(global_shape=self.shape[1] if adjoint else self.shape[0],
                             base_comm=x.base_comm, partition=x.partition, dtype=self._tdtype)
        xl = x.local_array
        if xl.dtype != self._tdtype:
            xl = xl.to(self._tdtype)
        per = nin * self.nz
        xs = xl.reshape(-1)[self.islstart[rank] * per: self.islend[rank] * per]
        multi = x.size > 1
        y1 = (torch.empty(self.nsl * nout * self.nz, dtype=self._tdtype, device=xl.device) if multi
              else y.local_array.view(-1))
        if self.nsl:
            _lib.check(_lib.lib.b2_batched_gemm(_lib.ctx(), self.G.data_ptr(), xs.data_ptr(), y1.data_ptr(),
                                                self.nsl, self.nx, self.ny, self.nz, int(adjoint),
                                                _lib.code(self._tdtype), _lib.stream()), "b2_batched_gemm")
        if multi:
            counts = [n * nout * self.nz for n in self.nsls]
            allgatherv(x.base_comm, y1, counts, out=y.local_array)
        return y

    def _matvec(self, x: DistributedArray) -> DistributedArray:
        return self._apply(x, False)

    def _rmatvec(self, x: DistributedArray) -> DistributedArray:
        return self._apply(x, True)

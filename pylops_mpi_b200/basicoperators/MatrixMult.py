"""``MPIMatrixMult`` -- distributed dense matrix product, "block" and "summa"
variants (pylops_mpi/basicoperators/MatrixMult.py:24-874).

Per-rank tile products run in libb200lops: ``b2_gemv`` when the local block has
a single column (HBM-bound), ``b2_gemm_bf16`` (tcgen05 tensor cores) for bf16
tiles with many columns, ``b2_gemm`` (SIMT) for float32/float64/complex
tiles.  Row / column sub-communicators are NCCL groups created once at
construction; the SUMMA adjoint keeps the reference's semantics but sources
each A^H tile from a transposed-grid copy exchanged ONCE at construction
(``saveAt``-style), turning the per-apply point-to-point tile routing of
MatrixMult.py:742-763 into a plain column broadcast.
"""
from __future__ import annotations

import math
from typing import Tuple

import numpy as np
import torch

from .. import _lib
from ..comm import COMM_WORLD, Comm, resolve, SUM
from ..Distributed import DistributedMixIn, allreduce_, allgatherv, bcast_, group, send, recv
from ..DistributedArray import DistributedArray, Partition
from ..LinearOperator import MPILinearOperator

__all__ = ["active_grid_comm", "block_gather", "local_block_split", "MPIMatrixMult"]


def active_grid_comm(base_comm, N: int, M: int):
    """MatrixMult.py:24-76: the square sub-grid of ranks that own data."""
    base_comm = resolve(base_comm)
    rank = base_comm.Get_rank()
    size = base_comm.Get_size()
    p_prime = math.isqrt(size)
    row, col = divmod(rank, p_prime)
    active_dim = min(N, M, p_prime)
    is_active = (row < active_dim and col < active_dim)
    # Split is collective: inactive ranks take part with a different colour
    new_comm = base_comm.Split(color=0 if is_active else 1, key=rank)
    if not is_active:
        return None, rank, row, col, False
    p_prime_new = math.isqrt(new_comm.Get_size())
    new_rank = new_comm.Get_rank()
    new_row, new_col = divmod(new_rank, p_prime_new)
    return new_comm, new_rank, new_row, new_col, True


def local_block_split(global_shape: Tuple[int, int], rank: int, comm) -> Tuple[slice, slice]:
    """MatrixMult.py:79-125 (integer bookkeeping, bit-exact)."""
    size = resolve(comm).Get_size()
    p_prime = math.isqrt(size)
    if p_prime * p_prime != size:
        raise RuntimeError(f"Number of processes must be a square number, "
                           f"provided {size} instead...")
    if not (isinstance(rank, int) and 0 <= rank < size):
        raise ValueError(f"rank must be an integer in [0, {size}), got {rank!r}")
    pr, pc = divmod(rank, p_prime)
    orig_r, orig_c = global_shape
    new_r = math.ceil(orig_r / p_prime) * p_prime
    new_c = math.ceil(orig_c / p_prime) * p_prime
    blkr, blkc = new_r // p_prime, new_c // p_prime
    rs, cs = pr * blkr, pc * blkc
    re, ce = min(rs + blkr, orig_r), min(cs + blkc, orig_c)
    return slice(rs, re), slice(cs, ce)


def block_gather(x: DistributedArray, orig_shape: Tuple[int, int], comm):
    """MatrixMult.py:128-175: assemble the 2-D block-distributed matrix on every rank."""
    comm = resolve(comm)
    p_prime = math.isqrt(comm.Get_size())
    if p_prime * p_prime != comm.Get_size():
        raise RuntimeError(f"Communicator size must be a perfect square, got {comm.Get_size()!r}")
    all_blks = x._allgather(comm, None, x.local_array)
    nr, nc = orig_shape
    br, bc = math.ceil(nr / p_prime), math.ceil(nc / p_prime)
    Cm = torch.zeros((nr, nc), dtype=all_blks[0].dtype, device=all_blks[0].device)
    for rank in range(p_prime * p_prime):
        pr, pc = divmod(rank, p_prime)
        rs, cs = pr * br, pc * bc
        re, ce = min(rs + br, nr), min(cs + bc, nc)
        if all_blks[rank].numel() != 0:
            Cm[rs:re, cs:ce] = all_blks[rank].reshape(re - rs, ce - cs)
    return Cm


def _to_device(A, dtype) -> torch.Tensor:
    if not isinstance(A, torch.Tensor):
        A = torch.as_tensor(np.asarray(A))
    return A.to(device="cuda", dtype=_lib.torch_dtype(dtype)).contiguous()


def _xdtype(adt: torch.dtype) -> torch.dtype:
    """dtype of the vectors an operator with matrix dtype ``adt`` works on"""
    return torch.float32 if adt is torch.bfloat16 else adt


def tile_product(A: torch.Tensor, X: torch.Tensor, Y: torch.Tensor, op: int, accumulate: bool):
    """Y (+)= op(A) X on the device.  A: 2-D tile; X: (k, ncol); Y: (m, ncol) contiguous."""
    m_a, n_a = A.shape
    m, k = (m_a, n_a) if op == _lib.OP_N else (n_a, m_a)
    ncol = X.shape[1]
    if m == 0 or ncol == 0:
        return Y
    if k == 0:
        if not accumulate:
            Y.zero_()
        return Y
    ctx, st = _lib.ctx(), _lib.stream()
    if A.dtype is torch.bfloat16:
        if ncol == 1:
            if accumulate:
                tmp = torch.empty_like(Y)
                _lib.check(_lib.lib.b2_gemv(ctx, A.data_ptr(), n_a, m_a, n_a, X.data_ptr(), tmp.data_ptr(),
                                            op, _lib.BF16, _lib.F32, st), "b2_gemv")
                one = _lib.cpair(1.0)
                _lib.check(_lib.lib.b2_lincomb(ctx, Y.data_ptr(), one, tmp.data_ptr(), one, Y.data_ptr(),
                                               Y.numel(), _lib.F32, 0, st), "b2_lincomb")
            else:
                _lib.check(_lib.lib.b2_gemv(ctx, A.data_ptr(), n_a, m_a, n_a, X.data_ptr(), Y.data_ptr(),
                                            op, _lib.BF16, _lib.F32, st), "b2_gemv")
            return Y
        Xb = X if X.dtype is torch.bfloat16 else X.to(torch.bfloat16)
        _lib.check(_lib.lib.b2_gemm_bf16(ctx, A.data_ptr(), n_a, Xb.data_ptr(), ncol, Y.data_ptr(), ncol,
                                         m, ncol, k, op, int(accumulate), st), "b2_gemm_bf16")
        return Y
    if ncol == 1 and not accumulate:
        _lib.check(_lib.lib.b2_gemv(ctx, A.data_ptr(), n_a, m_a, n_a, X.data_ptr(), Y.data_ptr(), op,
                                    _lib.code(A.dtype), _lib.code(X.dtype), st), "b2_gemv")
        return Y
    _lib.check(_lib.lib.b2_gemm(ctx, A.data_ptr(), n_a, X.data_ptr(), ncol, Y.data_ptr(), ncol, m, ncol, k,
                                op, int(accumulate), _lib.code(A.dtype), st), "b2_gemm")
    return Y


class _MPIBlockMatrixMult(DistributedMixIn, MPILinearOperator):
    """1-D block variant (MatrixMult.py:178-428): A split in row blocks over the
    grid columns, X in column blocks over the grid rows."""

    def __init__(self, A, M: int, saveAt: bool = False, base_comm=COMM_WORLD, dtype="float64",
                 base_comm_nccl=None) -> None:
        base_comm = resolve(base_comm)
        rank, size = base_comm.Get_rank(), base_comm.Get_size()
        self._P_prime = math.isqrt(size)
        self._C = self._P_prime
        if self._P_prime * self._C != size:
            raise Exception(f"Number of processes must be a square number, provided {size} instead...")
        self._col_id = rank % self._P_prime
        self._row_id = rank // self._P_prime
        self.base_comm = base_comm
        self._row_comm = base_comm.Split(color=self._row_id, key=self._col_id)
        self._col_comm = base_comm.Split(color=self._col_id, key=self._row_id)
        self.A = _to_device(A, dtype)
        if saveAt:
            self.At = self.A.T.conj().contiguous()
        rows = self._row_comm.allgather(int(self.A.shape[0]))
        self._row_counts = rows
        self.N = int(sum(rows))
        self.K = int(self.A.shape[1])
        self.M = int(M)
        block_cols = int(math.ceil(self.M / self._P_prime))
        blk_rows = int(math.ceil(self.N / self._P_prime))
        self._row_start = self._col_id * blk_rows
        self._row_end = min(self.N, self._row_start + blk_rows)
        self._col_start = self._row_id * block_cols
        self._col_end = min(self.M, self._col_start + block_cols)
        self._local_ncols = max(0, self._col_end - self._col_start)
        self._rank_col_lens = base_comm.allgather(self._local_ncols)
        total_ncols = int(np.sum(self._rank_col_lens))
        self.dims = (self.K, total_ncols)
        self.dimsd = (self.N, total_ncols)
        shape = (int(np.prod(self.dimsd)), int(np.prod(self.dims)))
        MPILinearOperator.__init__(self, shape=shape, dtype=_lib.numpy_dtype(_xdtype(self.A.dtype)),
                                   base_comm=base_comm)

    def _matvec(self, x: DistributedArray) -> DistributedArray:
        if x.partition != Partition.SCATTER:
            raise ValueError(f"x should have partition={Partition.SCATTER} Got {x.partition} instead...")
        xdt = _xdtype(self.A.dtype)
        y = DistributedArray(global_shape=(self.N * self.dimsd[1]),
                             local_shapes=[(self.N * c) for c in self._rank_col_lens],
                             mask=x.mask, partition=Partition.SCATTER, dtype=xdt, base_comm=x.base_comm)
        nc = self._rank_col_lens[self.rank]
        X = x.local_array.to(xdt).reshape(self.dims[0], nc)
        Yloc = torch.empty((self.A.shape[0], nc), dtype=xdt, device=X.device)
        tile_product(self.A, X.contiguous(), Yloc, _lib.OP_N, False)
        # row-comm Allgather + vstack == concatenation of row blocks (MatrixMult.py:370-377)
        counts = [r * nc for r in self._row_counts]
        allgatherv(self._row_comm, Yloc.view(-1), counts, out=y.local_array)
        return y

    def _rmatvec(self, x: DistributedArray) -> DistributedArray:
        if x.partition != Partition.SCATTER:
            raise ValueError(f"x should have partition={Partition.SCATTER}. Got {x.partition} instead.")
        xdt = _xdtype(self.A.dtype)
        y = DistributedArray(global_shape=(self.K * self.dimsd[1]),
                             local_shapes=[self.K * c for c in self._rank_col_lens],
                             mask=x.mask, partition=Partition.SCATTER, dtype=xdt, base_comm=x.base_comm)
        nc = self._local_ncols
        X = x.local_array.to(xdt).reshape(self.N, nc)
        X_tile = X[self._row_start:self._row_end, :].contiguous()
        Yloc = y.local_array.view(self.K, nc)
        if hasattr(self, "At"):
            tile_product(self.At, X_tile, Yloc, _lib.OP_N, False)
        else:
            tile_product(self.A, X_tile, Yloc, _lib.OP_H, False)
        allreduce_(self._row_comm, y.local_array, SUM)     # MatrixMult.py:420-426
        return y


class _MPISummaMatrixMult(DistributedMixIn, MPILinearOperator):
    """2-D SUMMA variant (MatrixMult.py:431-767) on a P' x P' grid."""

    def __init__(self, A, M: int, saveAt: bool = False, base_comm=COMM_WORLD, dtype="float64",
                 base_comm_nccl=None) -> None:
        base_comm = resolve(base_comm)
        rank, size = base_comm.Get_rank(), base_comm.Get_size()
        self._P_prime = math.isqrt(size)
        if self._P_prime * self._P_prime != size:
            raise Exception(f"Number of processes must be a square number, provided {size} instead...")
        P = self._P_prime
        self._row_id, self._col_id = divmod(rank, P)
        self.base_comm = base_comm
        self._row_comm = base_comm.Split(color=self._row_id, key=self._col_id)
        self._col_comm = base_comm.Split(color=self._col_id, key=self._row_id)
        A = _to_device(A, dtype)
        self.N = int(self._col_comm.allreduce(int(A.shape[0])))
        self.K = int(self._row_comm.allreduce(int(A.shape[1])))
        self.M = int(M)
        self._N_padded = math.ceil(self.N / P) * P
        self._K_padded = math.ceil(self.K / P) * P
        self._M_padded = math.ceil(self.M / P) * P
        bn, bk = self._N_padded // P, self._K_padded // P
        pr = (bn - A.shape[0]) if self._row_id == P - 1 else 0
        pc = (bk - A.shape[1]) if self._col_id == P - 1 else 0
        if pr > 0 or pc > 0:
            A = torch.nn.functional.pad(A, (0, pc, 0, pr))
        self.A = A.contiguous()
        # transposed-grid copy of the tile: rank (i, j) keeps A_{j,i} so that the adjoint is a
        # standard SUMMA with A^H row-broadcasts replaced by column-local products
        self._At_src = self._exchange_transposed(self.A)
        self.dims = (self.K, self.M)
        self.dimsd = (self.N, self.M)
        shape = (int(np.prod(self.dimsd)), int(np.prod(self.dims)))
        MPILinearOperator.__init__(self, shape=shape, dtype=_lib.numpy_dtype(_xdtype(self.A.dtype)),
                                   base_comm=base_comm)

    def _exchange_transposed(self, A: torch.Tensor) -> torch.Tensor:
        """tile A_{col_id,row_id} (from the transposed grid position), exchanged once"""
        P = self._P_prime
        partner = self._col_id * P + self._row_id
        if partner == self.base_comm.Get_rank() or self.base_comm.Get_size() == 1:
            return A
        out = torch.empty_like(A)
        with group(self.base_comm):
            send(self.base_comm, A, partner)
            recv(self.base_comm, out, partner)
        return out

    def _local_extent(self, b: int, full: int, idx: int) -> int:
        return b if idx != self._P_prime - 1 else full - (self._P_prime - 1) * b

    def _padded_block(self, x: DistributedArray, rows_b: int, rows_full: int, xdt) -> Tuple[torch.Tensor, int, int]:
        P = self._P_prime
        bm = self._M_padded // P
        local_r = self._local_extent(rows_b, rows_full, self._row_id)
        local_m = self._local_extent(bm, self.M, self._col_id)
        blk = x.local_array.to(xdt).reshape(local_r, local_m)
        if local_r != rows_b or local_m != bm:
            blk = torch.nn.functional.pad(blk, (0, bm - local_m, 0, rows_b - local_r))
        return blk.contiguous(), local_r, local_m

    def _matvec(self, x: DistributedArray) -> DistributedArray:
        if x.partition != Partition.SCATTER:
            raise ValueError(f"x should have partition={Partition.SCATTER} Got {x.partition} instead...")
        P = self._P_prime
        xdt = _xdtype(self.A.dtype)
        bn, bk, bm = self._N_padded // P, self._K_padded // P, self._M_padded // P
        local_n = self._local_extent(bn, self.N, self._row_id)
        sizes = []
        for r in range(self.size):
            ri, ci = divmod(r, P)
            sizes.append(self._local_extent(bn, self.N, ri) * self._local_extent(bm, self.M, ci))
        y = DistributedArray(global_shape=(self.N * self.M), mask=x.mask, local_shapes=sizes,
                             partition=Partition.SCATTER, dtype=xdt, base_comm=x.base_comm)
        x_block, local_k, local_m = self._padded_block(x, bk, self.K, xdt)
        Y_local = torch.empty((self.A.shape[0], bm), dtype=xdt, device=x_block.device)
        Atemp = torch.empty_like(self.A) if P > 1 else None
        Xtemp = torch.empty_like(x_block) if P > 1 else None
        for k in range(P):
            # MatrixMult.py:663-670: Bcast A_k along the row, X_k along the column, accumulate
            if P == 1:
                a_k, x_k = self.A, x_block
            else:
                a_k = self.A if self._col_id == k else Atemp
                x_k = x_block if self._row_id == k else Xtemp
                bcast_(self._row_comm, a_k, root=k)
                bcast_(self._col_comm, x_k, root=k)
            tile_product(a_k, x_k, Y_local, _lib.OP_N, accumulate=(k > 0))
        y.local_array.copy_(Y_local[:local_n, :local_m].reshape(-1))
        return y

    def _rmatvec(self, x: DistributedArray) -> DistributedArray:
        if x.partition != Partition.SCATTER:
            raise ValueError(f"x should have partition={Partition.SCATTER}. Got {x.partition} instead.")
        P = self._P_prime
        xdt = _xdtype(self.A.dtype)
        bn, bk, bm = self._N_padded // P, self._K_padded // P, self._M_padded // P
        local_k = self._local_extent(bk, self.K, self._row_id)
        sizes = []
        for r in range(self.size):
            ri, ci = divmod(r, P)
            sizes.append(self._local_extent(bk, self.K, ri) * self._local_extent(bm, self.M, ci))
        y = DistributedArray(global_shape=(self.K * self.M), mask=x.mask, local_shapes=sizes,
                             partition=Partition.SCATTER, dtype=xdt, base_comm=x.base_comm)
        x_block, local_n, local_m = self._padded_block(x, bn, self.N, xdt)
        Y_local = torch.empty((self.A.shape[1], bm), dtype=xdt, device=x_block.device)
        # result tile (i, j) = sum_k (A_{k,i})^H X_{k,j}  (MatrixMult.py:742-763).  Rank (i, j)
        # holds A_{j,i} (transposed-grid copy); A_{k,i} lives on rank (i, k) of grid row i ->
        # a ROW broadcast of the transposed copies, root k.
        Atemp = torch.empty_like(self._At_src) if P > 1 else None
        Xtemp = torch.empty_like(x_block) if P > 1 else None
        for k in range(P):
            if P == 1:
                a_k, x_k = self._At_src, x_block
            else:
                a_k = self._At_src if self._col_id == k else Atemp
                x_k = x_block if self._row_id == k else Xtemp
                bcast_(self._row_comm, a_k, root=k)
                bcast_(self._col_comm, x_k, root=k)
            tile_product(a_k, x_k, Y_local, _lib.OP_H, accumulate=(k > 0))
        y.local_array.copy_(Y_local[:local_k, :local_m].reshape(-1))
        return y


def MPIMatrixMult(A, M: int, saveAt: bool = False, base_comm=COMM_WORLD, kind: str = "summa",
                  dtype="float64", base_comm_nccl=None):
    """Factory with the reference's signature (MatrixMult.py:770-874)."""
    if kind == "summa":
        return _MPISummaMatrixMult(A, M, saveAt, base_comm, dtype, base_comm_nccl)
    elif kind == "block":
        return _MPIBlockMatrixMult(A, M, saveAt, base_comm, dtype, base_comm_nccl)
    else:
        raise NotImplementedError("kind must be summa or block")

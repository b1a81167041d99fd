"""CPU ORACLE -- TEST INFRASTRUCTURE ONLY (never imported by the product path).

NumPy restatement of the pylops-mpi distributed matvec/rmatvec hot path,
simulating P MPI ranks inside one process.  Every "distributed array" here is
a python list with one NumPy array per simulated rank; every collective is the
obvious list operation (Allreduce = sum over the list in rank order,
Allgather = the list itself, Send/Recv = copying a slice from the neighbour's
entry).  Each function cites the reference file:line (relative to
/root/reference/pylops_mpi/) whose per-rank algorithm it follows.

Pinning status: this restatement is checked (tests/test_oracle.py) against
  * every known-answer vector the reference's own tests hold for this path that
    is computable with NumPy alone, and
  * fixtures produced by the *real* reference code (imported from
    /root/reference under an in-process MPI shim, see
    tests/golden/make_golden.py) committed under tests/golden/.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module.
"""
from __future__ import annotations

import math
from typing import Callable, List, Optional, Sequence, Tuple

import numpy as np

BROADCAST = "Broadcast"
UNSAFE_BROADCAST = "UnsafeBroadcast"
SCATTER = "Scatter"


# --------------------------------------------------------------------------
# partition bookkeeping (integer, bit-exact)
# --------------------------------------------------------------------------
def local_split(global_shape: Tuple[int, ...], size: int, rank: int,
                partition: str = SCATTER, axis: int = 0) -> Tuple[int, ...]:
    """DistributedArray.py:42-71"""
    if partition in (BROADCAST, UNSAFE_BROADCAST):
        return tuple(global_shape)
    local_shape = list(global_shape)
    if rank < (global_shape[axis] % size):
        local_shape[axis] = global_shape[axis] // size + 1
    else:
        local_shape[axis] = global_shape[axis] // size
    return tuple(local_shape)


def local_shapes(global_shape, size, partition=SCATTER, axis=0):
    """DistributedArray.py:345-358 (allgather of local_split)."""
    return [local_split(tuple(global_shape), size, r, partition, axis) for r in range(size)]


def to_dist(x: np.ndarray, size: int, partition: str = SCATTER, axis: int = 0,
            shapes: Optional[Sequence[Tuple[int, ...]]] = None) -> List[np.ndarray]:
    """DistributedArray.py:407-460: every rank slices its own block of the
    (replicated) global array at offsets cumsum(local extents)."""
    if partition in (BROADCAST, UNSAFE_BROADCAST):
        return [x.copy() for _ in range(size)]
    if shapes is None:
        shapes = local_shapes(x.shape, size, partition, axis)
    ext = np.append([0], [s[axis] for s in shapes])
    off = np.cumsum(ext)
    out = []
    for r in range(size):
        sl = [slice(None)] * x.ndim
        sl[axis] = slice(off[r], off[r + 1])
        out.append(x[tuple(sl)].copy())
    return out


def asarray(locs: List[np.ndarray], partition: str = SCATTER, axis: int = 0) -> np.ndarray:
    """DistributedArray.py:370-405"""
    if partition in (BROADCAST, UNSAFE_BROADCAST):
        return locs[0]
    return np.concatenate(locs, axis=axis)


# --------------------------------------------------------------------------
# reductions
# --------------------------------------------------------------------------
def dot(x: List[np.ndarray], y: List[np.ndarray], vdot: bool = False,
        partition: str = SCATTER, mask: Optional[Sequence[int]] = None):
    """DistributedArray.py:654-686.  BROADCAST operands are first re-scattered
    (:678-681); the Allreduce runs on the mask sub-communicator, so the result
    is a list with one value per rank."""
    size = len(x)
    if partition in (BROADCAST, UNSAFE_BROADCAST):
        x = to_dist(x[0], size)
        y = to_dist(y[0], size)
    f = np.vdot if vdot else np.dot
    part = [f(x[r].flatten(), y[r].flatten()) for r in range(size)]
    return _allreduce_sub(part, mask, "sum")


def _allreduce_sub(part, mask, op):
    size = len(part)
    mask = [0] * size if mask is None else list(mask)
    out = [None] * size
    for color in sorted(set(mask)):
        members = [r for r in range(size) if mask[r] == color]
        vals = [part[r] for r in members]
        if op == "sum":
            red = vals[0]
            for v in vals[1:]:
                red = red + v
        elif op == "max":
            red = np.maximum.reduce(vals)
        else:
            red = np.minimum.reduce(vals)
        for r in members:
            out[r] = red
    return out


def norm(x: List[np.ndarray], ord=None, partition: str = SCATTER,
         mask: Optional[Sequence[int]] = None):
    """DistributedArray.py:688-758 + :774-807 for axis=None (flattened)."""
    size = len(x)
    if partition in (BROADCAST, UNSAFE_BROADCAST):
        x = to_dist(x[0], size)
    ord = 2 if ord is None else ord
    flat = [a.flatten() for a in x]
    if ord in ("fro", "nuc"):
        raise ValueError(f"norm-{ord} not possible for vectors")
    if ord == 0:
        part = [np.float64(np.count_nonzero(a)) for a in flat]
        return _allreduce_sub(part, mask, "sum")
    if ord == np.inf:
        part = [np.float64(np.max(np.abs(a))) if a.size else np.float64(0) for a in flat]
        return _allreduce_sub(part, mask, "max")
    if ord == -np.inf:
        part = [np.float64(np.min(np.abs(a))) if a.size else np.float64(np.inf) for a in flat]
        return _allreduce_sub(part, mask, "min")
    part = [np.sum(np.abs(np.float_power(a, ord))) for a in flat]  # :755 float64 promote
    red = _allreduce_sub(part, mask, "sum")
    return [np.power(v, 1.0 / ord) for v in red]


# --------------------------------------------------------------------------
# halo exchange and @reshaped redistribution
# --------------------------------------------------------------------------
def add_ghost_cells(locs: List[np.ndarray], axis: int = 0,
                    cells_front: Optional[Sequence[int]] = None,
                    cells_back: Optional[Sequence[int]] = None) -> List[np.ndarray]:
    """DistributedArray.py:876-953.  cells_front/cells_back are per-rank lists
    (what each rank passes); tag-1 messages flow rank -> rank+1, tag-0 flow
    rank -> rank-1.  Raises the reference's ValueError when a rank is asked
    for more cells than it owns."""
    size = len(locs)
    ghosted = [a.copy() for a in locs]
    if cells_front is not None:
        total = list(cells_front) + [0]
        for rank in range(size):
            want = total[rank + 1]  # what rank+1 needs from me
            send_buf = np.take(locs[rank], np.arange(-want, 0), axis=axis)
            if rank != 0:
                if total[rank] != 0 and locs[rank - 1].shape[axis] != 0:
                    need = total[rank]
                    if need > locs[rank - 1].shape[axis]:
                        raise ValueError(f"Local Shape at rank={rank - 1} along axis={axis} "
                                         f"should be > {need}")
                    recv = np.take(locs[rank - 1], np.arange(-need, 0), axis=axis)
                    ghosted[rank] = np.concatenate([recv, ghosted[rank]], axis=axis)
            if rank != size - 1 and len(send_buf) != 0:
                if want > locs[rank].shape[axis]:
                    raise ValueError(f"Local Shape at rank={rank} along axis={axis} "
                                     f"should be > {want}")
    if cells_back is not None:
        total = list(cells_back) + [0]
        for rank in range(size):
            want = total[rank - 1]  # what rank-1 needs from me (wraps for rank 0)
            send_buf = np.take(locs[rank], np.arange(want), axis=axis)
            if rank != 0 and len(send_buf) != 0:
                if want > locs[rank].shape[axis]:
                    raise ValueError(f"Local Shape at rank={rank} along axis={axis} "
                                     f"should be > {want}")
            if rank != size - 1:
                need = total[rank]
                if need != 0 and locs[rank + 1].shape[axis] != 0:
                    if need > locs[rank + 1].shape[axis]:
                        raise ValueError(f"Local Shape at rank={rank + 1} along axis={axis} "
                                         f"should be > {need}")
                    recv = np.take(locs[rank + 1], np.arange(need), axis=axis)
                    ghosted[rank] = np.append(ghosted[rank], recv, axis=axis)
    return ghosted


def reshaped_in(x: List[np.ndarray], arr_local_shapes: Sequence[Tuple[int, ...]]) -> List[np.ndarray]:
    """utils/decorators.py:44-72: re-partition a flat SCATTER vector to the
    operator's row-block partition using neighbour ghost cells only."""
    size = len(x)
    arr_sizes = np.asarray([int(np.prod(s)) for s in arr_local_shapes])
    x_sizes = np.asarray([int(np.prod(a.shape)) for a in x])
    dif = np.cumsum(arr_sizes - x_sizes)
    cells_front = [abs(min(0, dif[r - 1])) for r in range(size)]
    cells_back = [max(0, dif[r]) for r in range(size)]
    ghosted = add_ghost_cells(x, 0, cells_front, cells_back)
    out = []
    for r in range(size):
        index = max(0, dif[r - 1])
        out.append(ghosted[r][index: arr_sizes[r] + index].reshape(arr_local_shapes[r]))
    return out


# --------------------------------------------------------------------------
# MPIFirstDerivative  (basicoperators/FirstDerivative.py)
# --------------------------------------------------------------------------
def _z(n, dims, dtype):
    return np.zeros((max(n, 0),) + tuple(dims[1:]), dtype=dtype)


def first_derivative(x_flat: List[np.ndarray], dims: Tuple[int, ...], sampling: float = 1.0,
                     kind: str = "centered", edge: bool = False, order: int = 3,
                     adjoint: bool = False, dtype=np.float64) -> List[np.ndarray]:
    """FirstDerivative.py:129-319 applied to a flat SCATTER vector (per-rank
    list) -> flat SCATTER vector with the dims-row-block local sizes."""
    size = len(x_flat)
    dims = tuple(int(d) for d in (dims if np.ndim(dims) else (dims,)))
    shapes = local_shapes(dims, size, SCATTER, 0)
    x = reshaped_in(x_flat, shapes)
    N = dims[0]
    gh = lambda **kw: add_ghost_cells(x, 0,  # noqa: E731
                                      [kw["front"]] * size if "front" in kw else None,
                                      [kw["back"]] * size if "back" in kw else None)
    y = [np.zeros(s, dtype=dtype) for s in shapes]
    last = size - 1
    if kind == "forward" and not adjoint:          # :141-151
        g = gh(back=1)
        for r in range(size):
            yf = g[r][1:] - g[r][:-1]
            if r == last:
                yf = np.append(yf, _z(1, dims, yf.dtype), axis=0)
            y[r][:] = yf / sampling
    elif kind == "forward" and adjoint:            # :153-169
        g = gh(front=1)
        for r in range(size):
            if r == last:
                y[r][:-1] -= x[r][:-1]
            else:
                y[r][:] -= x[r][:]
            yf = g[r][:-1]
            if r == 0:
                yf = np.append(_z(1, dims, yf.dtype), yf, axis=0)
            y[r][:] += yf
            y[r][:] /= sampling
    elif kind == "backward" and not adjoint:       # :171-181
        g = gh(front=1)
        for r in range(size):
            yb = g[r][1:] - g[r][:-1]
            if r == 0:
                yb = np.append(_z(1, dims, yb.dtype), yb, axis=0)
            y[r][:] = yb / sampling
    elif kind == "backward" and adjoint:           # :183-199
        g = gh(back=1)
        for r in range(size):
            yb = g[r][1:]
            if r == last:
                yb = np.append(yb, _z(1, dims, yb.dtype), axis=0)
            y[r][:] -= yb
            if r == 0:
                y[r][1:] += x[r][1:]
            else:
                y[r][:] += x[r][:]
            y[r][:] /= sampling
    elif kind == "centered" and order == 3 and not adjoint:   # :201-219
        g = gh(front=1, back=1)
        for r in range(size):
            yc = 0.5 * (g[r][2:] - g[r][:-2])
            if r == 0:
                yc = np.append(_z(1, dims, yc.dtype), yc, axis=0)
            if r == last:
                yc = np.append(yc, _z(min(N - 1, 1), dims, yc.dtype), axis=0)
            y[r][:] = yc
            if edge:
                if r == 0:
                    y[r][0] = x[r][1] - x[r][0]
                if r == last:
                    y[r][-1] = x[r][-1] - x[r][-2]
            y[r][:] /= sampling
    elif kind == "centered" and order == 3 and adjoint:       # :221-247
        g1 = gh(back=2)
        g2 = gh(front=2)
        for r in range(size):
            yc = 0.5 * g1[r][1:-1]
            if r == last:
                yc = np.append(yc, _z(min(N, 2), dims, yc.dtype), axis=0)
            y[r][:] -= yc
            yc = 0.5 * g2[r][1:-1]
            if r == 0:
                yc = np.append(_z(min(N, 2), dims, yc.dtype), yc, axis=0)
            y[r][:] += yc
            if edge:
                if r == 0:
                    y[r][0] -= x[r][0]
                    y[r][1] += x[r][0]
                if r == last:
                    y[r][-2] -= x[r][-1]
                    y[r][-1] += x[r][-1]
            y[r][:] /= sampling
    elif kind == "centered" and order == 5 and not adjoint:   # :249-274
        g = gh(front=2, back=2)
        for r in range(size):
            yc = (g[r][:-4] / 12.0 - 2 * g[r][1:-3] / 3.0
                  + 2 * g[r][3:-1] / 3.0 - g[r][4:] / 12.0)
            if r == 0:
                yc = np.append(_z(min(N, 2), dims, yc.dtype), yc, axis=0)
            if r == last:
                yc = np.append(yc, _z(min(N - 2, 2), dims, yc.dtype), axis=0)
            y[r][:] = yc
            if edge:
                if r == 0:
                    y[r][0] = x[r][1] - x[r][0]
                    y[r][1] = 0.5 * (x[r][2] - x[r][0])
                if r == last:
                    y[r][-1] = x[r][-1] - x[r][-2]
                    y[r][-2] = 0.5 * (x[r][-1] - x[r][-3])
            y[r][:] /= sampling
    elif kind == "centered" and order == 5 and adjoint:       # :276-319
        ga = gh(back=4)
        gb = add_ghost_cells(x, 0, [1] * size, [3] * size)
        gc = add_ghost_cells(x, 0, [3] * size, [1] * size)
        gd = gh(front=4)
        for r in range(size):
            yc = ga[r][2:-2] / 12.0
            if r == last:
                yc = np.append(yc, _z(min(N, 4), dims, yc.dtype), axis=0)
            y[r][:] += yc
            yc = 2.0 * gb[r][2:-2] / 3.0
            if r == 0:
                yc = np.append(_z(1, dims, yc.dtype), yc, axis=0)
            if r == last:
                yc = np.append(yc, _z(min(N - 1, 3), dims, yc.dtype), axis=0)
            y[r][:] -= yc
            yc = 2.0 * gc[r][2:-2] / 3.0
            if r == 0:
                yc = np.append(_z(min(N, 3), dims, yc.dtype), yc, axis=0)
            if r == last:
                yc = np.append(yc, _z(min(N - 3, 1), dims, yc.dtype), axis=0)
            y[r][:] += yc
            yc = gd[r][2:-2] / 12.0
            if r == 0:
                yc = np.append(_z(min(N, 4), dims, yc.dtype), yc, axis=0)
            y[r][:] -= yc
            if edge:
                if r == 0:
                    y[r][0] -= x[r][0] + 0.5 * x[r][1]
                    y[r][1] += x[r][0]
                    y[r][2] += 0.5 * x[r][1]
                if r == last:
                    y[r][-3] -= 0.5 * x[r][-2]
                    y[r][-2] -= x[r][-1]
                    y[r][-1] += 0.5 * x[r][-2] + x[r][-1]
            y[r][:] /= sampling
    else:
        raise NotImplementedError("'kind' must be 'forward', 'centered', or 'backward'; "
                                  "'order' must be '3, or '5'")
    return [a.ravel() for a in y]   # decorators.py:74-75


def first_derivative_dense(N: int, sampling=1.0, kind="centered", edge=False, order=3) -> np.ndarray:
    """Serial N x N stencil matrix (the operator pylops.FirstDerivative applies
    along axis 0; the reference tests compare against it,
    tests/test_derivative.py:220-229).  Independent of the per-rank code above;
    used to cross-check it."""
    D = np.zeros((N, N))
    if kind == "forward":
        for i in range(N - 1):
            D[i, i], D[i, i + 1] = -1, 1
    elif kind == "backward":
        for i in range(1, N):
            D[i, i - 1], D[i, i] = -1, 1
    elif kind == "centered" and order == 3:
        for i in range(1, N - 1):
            D[i, i - 1], D[i, i + 1] = -0.5, 0.5
        if edge:
            D[0, 0], D[0, 1] = -1, 1
            D[N - 1, N - 2], D[N - 1, N - 1] = -1, 1
    elif kind == "centered" and order == 5:
        for i in range(2, N - 2):
            D[i, i - 2], D[i, i - 1], D[i, i + 1], D[i, i + 2] = 1 / 12, -2 / 3, 2 / 3, -1 / 12
        if edge:
            D[0, 0], D[0, 1] = -1, 1
            D[1, 0], D[1, 2] = -0.5, 0.5
            D[N - 2, N - 3], D[N - 2, N - 1] = -0.5, 0.5
            D[N - 1, N - 2], D[N - 1, N - 1] = -1, 1
    else:
        raise NotImplementedError
    return D / sampling


# --------------------------------------------------------------------------
# MPIBlockDiag / MPIVStack  (blocks are dense matrices = pylops.MatrixMult)
# --------------------------------------------------------------------------
def blockdiag(blocks: List[List[np.ndarray]], x: List[np.ndarray], adjoint=False) -> List[np.ndarray]:
    """BlockDiag.py:121-143.  blocks[r] = list of dense matrices owned by rank
    r; x is a flat SCATTER vector (any per-rank split) -> re-partitioned to the
    operator's local_shapes_m / local_shapes_n first (decorators.py:47-52)."""
    size = len(blocks)
    if adjoint:
        want = [(sum(b.shape[0] for b in blocks[r]),) for r in range(size)]
    else:
        want = [(sum(b.shape[1] for b in blocks[r]),) for r in range(size)]
    xr = reshaped_in(x, want)
    out = []
    for r in range(size):
        off, y1 = 0, []
        for b in blocks[r]:
            if adjoint:
                y1.append(b.conj().T @ xr[r][off: off + b.shape[0]])
                off += b.shape[0]
            else:
                y1.append(b @ xr[r][off: off + b.shape[1]])
                off += b.shape[1]
        out.append(np.concatenate(y1))
    return out


def vstack_matvec(blocks: List[List[np.ndarray]], x_bcast: np.ndarray) -> List[np.ndarray]:
    """VStack.py:120-132 (x BROADCAST -> y SCATTER)."""
    return [np.concatenate([b @ x_bcast for b in blocks[r]]) for r in range(len(blocks))]


def vstack_rmatvec(blocks: List[List[np.ndarray]], x: List[np.ndarray]) -> np.ndarray:
    """VStack.py:134-149 (x SCATTER -> y BROADCAST via Allreduce SUM)."""
    size = len(blocks)
    want = [(sum(b.shape[0] for b in blocks[r]),) for r in range(size)]
    xr = reshaped_in(x, want)
    part = []
    for r in range(size):
        off, y1 = 0, []
        for b in blocks[r]:
            y1.append(b.conj().T @ xr[r][off: off + b.shape[0]])
            off += b.shape[0]
        part.append(np.sum(np.vstack(y1), axis=0))
    red = part[0]
    for p in part[1:]:
        red = red + p
    return red


# --------------------------------------------------------------------------
# MPIMatrixMult  (basicoperators/MatrixMult.py)
# --------------------------------------------------------------------------
def local_block_split(global_shape: Tuple[int, int], rank: int, size: int):
    """MatrixMult.py:79-125"""
    p_prime = math.isqrt(size)
    if p_prime * p_prime != size:
        raise RuntimeError(f"Number of processes must be a square number, provided {size} instead...")
    pr, pc = divmod(rank, p_prime)
    orig_r, orig_c = global_shape
    new_r = math.ceil(orig_r / p_prime) * p_prime
    new_c = math.ceil(orig_c / p_prime) * p_prime
    blkr, blkc = new_r // p_prime, new_c // p_prime
    rs, cs = pr * blkr, pc * blkc
    re, ce = min(rs + blkr, orig_r), min(cs + blkc, orig_c)
    return slice(rs, re), slice(cs, ce)


def block_gather(locs: List[np.ndarray], orig_shape: Tuple[int, int]) -> np.ndarray:
    """MatrixMult.py:128-175 (with the (re-rs, ce-cs) extent)."""
    size = len(locs)
    p_prime = math.isqrt(size)
    nr, nc = orig_shape
    br, bc = math.ceil(nr / p_prime), math.ceil(nc / p_prime)
    C = np.zeros((nr, nc), dtype=locs[0].dtype)
    for rank in range(size):
        pr, pc = divmod(rank, p_prime)
        rs, cs = pr * br, pc * bc
        re, ce = min(rs + br, nr), min(cs + bc, nc)
        if len(locs[rank]) != 0:
            C[rs:re, cs:ce] = locs[rank].reshape(re - rs, ce - cs)
    return C


def summa_tiles(A: np.ndarray, size: int) -> List[np.ndarray]:
    """How the reference tests cut a global matrix into the 2-D tiles each rank
    owns (tests/test_matrixmult.py:108-117 via local_block_split)."""
    return [A[local_block_split(A.shape, r, size)].copy() for r in range(size)]


def summa_matvec(A_tiles: List[np.ndarray], x: List[np.ndarray], N: int, K: int, M: int,
                 dtype=np.float64, adjoint: bool = False) -> List[np.ndarray]:
    """_MPISummaMatrixMult._matvec / _rmatvec, MatrixMult.py:612-767, with the
    broadcasts / p2p tile routing replaced by list lookups.
    x[r] is the flat (local_k x local_m) [fwd] or (local_n x local_m) [adj] tile."""
    size = len(A_tiles)
    P = math.isqrt(size)
    if P * P != size:
        raise Exception(f"Number of processes must be a square number, provided {size} instead...")
    Np, Kp, Mp = (math.ceil(v / P) * P for v in (N, K, M))
    bn, bk, bm = Np // P, Kp // P, Mp // P
    Apad = []
    for r in range(size):
        row, col = divmod(r, P)
        a = A_tiles[r].astype(dtype)
        pr = (bn - a.shape[0]) if row == P - 1 else 0
        pc = (bk - a.shape[1]) if col == P - 1 else 0
        if pr > 0 or pc > 0:
            a = np.pad(a, [(0, pr), (0, pc)], mode="constant")
        Apad.append(a)
    out = []
    xdt = x[0].dtype
    if not adjoint:
        out_dtype = np.result_type(np.dtype(dtype), xdt)
    elif np.iscomplexobj(Apad[0]):
        out_dtype = np.result_type(np.dtype(dtype), xdt)
    else:
        out_dtype = np.result_type(np.dtype(dtype), xdt if np.issubdtype(xdt, np.complexfloating) else np.dtype(dtype))
    acc = out_dtype
    if np.issubdtype(out_dtype, np.complexfloating):
        acc = np.promote_types(out_dtype, np.longdouble if out_dtype == np.complex128 else np.float64)

    def loc(b, full, idx):
        return b if idx != P - 1 else full - (P - 1) * b

    xb = []
    for r in range(size):
        row, col = divmod(r, P)
        rows = loc(bn if adjoint else bk, N if adjoint else K, row)
        lm = loc(bm, M, col)
        blk = x[r].reshape((rows, lm)).astype(acc)
        pad_r = (bn if adjoint else bk) - rows
        pad_m = bm - lm
        if pad_r > 0 or pad_m > 0:
            blk = np.pad(blk, [(0, pad_r), (0, pad_m)], mode="constant")
        xb.append(blk)
    for r in range(size):
        row, col = divmod(r, P)
        if not adjoint:
            Y = np.zeros((Apad[r].shape[0], bm), dtype=acc)
            for k in range(P):
                Atemp = Apad[row * P + k].astype(acc)      # Bcast on row-comm, root k
                Xtemp = xb[k * P + col]                    # Bcast on col-comm, root k
                Y += np.dot(Atemp, Xtemp)
            rows_out = loc(bn, N, row)
        else:
            Y = np.zeros((Apad[r].shape[1], bm), dtype=acc)
            for k in range(P):
                Xtemp = xb[k * P + col]                    # Bcast on col-comm, root k
                srcA = k * P + row                         # MatrixMult.py:748
                ATtemp = Apad[srcA].T.conj().astype(acc)
                Y += np.dot(ATtemp, Xtemp)
            rows_out = loc(bk, K, row)
        lm = loc(bm, M, col)
        out.append(Y[:rows_out, :lm].astype(out_dtype).flatten())
    return out


def blockmm_matvec(A_rows: List[np.ndarray], x: List[np.ndarray], N: int, K: int, M: int,
                   dtype=np.float64, adjoint: bool = False) -> List[np.ndarray]:
    """_MPIBlockMatrixMult, MatrixMult.py:280-428.  A_rows[r] is the row block
    rank r was constructed with (selected by col_id = r % P'); x[r] flat
    (K x ncols_r) [fwd] or (N x ncols_r) [adj]."""
    size = len(A_rows)
    P = math.isqrt(size)
    if P * P != size:
        raise Exception(f"Number of processes must be a square number, provided {size} instead...")
    block_cols = int(math.ceil(M / P))
    blk_rows = int(math.ceil(N / P))
    ncols = []
    for r in range(size):
        row_id = r // P
        cs = row_id * block_cols
        ce = min(M, cs + block_cols)
        ncols.append(max(0, ce - cs))
    out = []
    for r in range(size):
        row_id, col_id = divmod(r, P)
        if not adjoint:
            tiles = []
            for c in range(P):            # Allgather over row-comm (same row_id, all col_id)
                q = row_id * P + c
                Xq = x[q].reshape((K, ncols[q]))
                tiles.append(np.matmul(A_rows[q].astype(dtype), Xq))
            out.append(np.vstack(tiles).flatten())
        else:
            acc = None
            for c in range(P):            # Allreduce over row-comm
                q = row_id * P + c
                rs = c * blk_rows
                re = min(N, rs + blk_rows)
                Xq = x[q].reshape((N, ncols[q]))[rs:re, :]
                Yq = np.matmul(A_rows[q].astype(dtype).T.conj(), Xq)
                acc = Yq if acc is None else acc + Yq
            out.append(acc.flatten())
    return out


# --------------------------------------------------------------------------
# MPIFredholm1  (signalprocessing/Fredholm1.py)
# --------------------------------------------------------------------------
def fredholm1(G_loc: List[np.ndarray], x_bcast: np.ndarray, nz: int, adjoint=False) -> np.ndarray:
    """Fredholm1.py:109-171.  G_loc[r]: (nsl_r, nx, ny); x BROADCAST flat ->
    y BROADCAST flat (Allgather + vstack + ravel)."""
    nsls = [g.shape[0] for g in G_loc]
    if 1 in nsls:
        raise NotImplementedError("All ranks must have at least 2 or more elements in the first dimension")
    nx, ny = G_loc[0].shape[1:]
    ntot = sum(nsls)
    start = np.insert(np.cumsum(nsls)[:-1], 0, 0)
    end = np.cumsum(nsls)
    dims_in = (ntot, nx, nz) if adjoint else (ntot, ny, nz)
    xs = x_bcast.reshape(dims_in)
    ys = []
    for r, G in enumerate(G_loc):
        xr = xs[start[r]:end[r]]
        if adjoint:
            ys.append(np.matmul(G.transpose((0, 2, 1)).conj(), xr))
        else:
            ys.append(np.matmul(G, xr))
    return np.vstack(ys).ravel()


# --------------------------------------------------------------------------
# dottest / cgls on simulated arrays
# --------------------------------------------------------------------------
class SimArray:
    """Per-rank list with the DistributedArray arithmetic the solvers use
    (DistributedArray.py:574-652)."""

    def __init__(self, locs, partition=SCATTER):
        self.locs = [np.asarray(a) for a in locs]
        self.partition = partition

    def copy(self):
        return SimArray([a.copy() for a in self.locs], self.partition)

    def conj(self):
        return SimArray([a.conj() for a in self.locs], self.partition)

    def __neg__(self):
        return SimArray([-a for a in self.locs], self.partition)

    def __add__(self, o):
        return SimArray([a + b for a, b in zip(self.locs, o.locs)], self.partition)

    def __sub__(self, o):
        return self + (-o)

    def __iadd__(self, o):
        self.locs = [a + b for a, b in zip(self.locs, o.locs)]
        return self

    def __isub__(self, o):
        return self.__iadd__(-o)

    def __mul__(self, s):
        if isinstance(s, SimArray):
            return SimArray([a * b for a, b in zip(self.locs, s.locs)], self.partition)
        return SimArray([a * s for a in self.locs], self.partition)

    __rmul__ = __mul__

    def dot(self, o):
        return dot(self.locs, o.locs, partition=self.partition)[0]

    def norm(self, ord=None):
        return norm(self.locs, ord, partition=self.partition)[0]

    def asarray(self):
        return asarray(self.locs, self.partition)


def dottest(matvec: Callable, rmatvec: Callable, u: SimArray, v: SimArray, rtol=1e-6, atol=1e-21):
    """utils/dottest.py:76-107"""
    y = matvec(u)
    x = rmatvec(v)
    yy = np.vdot(y.asarray(), v.asarray())
    xx = np.vdot(u.asarray(), x.asarray())
    return bool(np.isclose(xx, yy, rtol, atol)), xx, yy


def cgls(matvec: Callable, rmatvec: Callable, y: SimArray, x0: SimArray, niter=10,
         damp=0.0, tol=1e-4):
    """optimization/cls_basic.py:308-404, :436, :451-469 (same recurrences,
    same damp / damp**2 quirk, same `kold > tol` stop rule)."""
    damp2 = damp ** 2
    x = x0.copy()
    s = y - matvec(x)
    r = rmatvec(s) - x * damp
    c = r.copy()
    q = matvec(c)
    kold = float(np.abs(r.dot(r.conj())))
    cost = [float(s.norm())]
    cost1 = [np.sqrt(float(cost[0] ** 2 + damp * np.abs(x.dot(x.conj()))))]
    iiter = 0
    while iiter < niter and kold > tol:
        a = float(np.abs(kold / (q.dot(q.conj()) + damp2 * c.dot(c.conj()))))
        x += a * c
        s -= a * q
        r = rmatvec(s) - damp2 * x
        k = float(np.abs(r.dot(r.conj())))
        b = float(k / kold)
        c = r + b * c
        q = matvec(c)
        kold = k
        iiter += 1
        cost.append(float(s.norm()))
        cost1.append(np.sqrt(float(cost[iiter] ** 2 + damp2 * np.abs(x.dot(x.conj())))))
    istop = 1 if kold < tol else 2
    return x, istop, iiter, kold, cost1[iiter], np.array(cost)


def cg(matvec: Callable, y: SimArray, x0: SimArray, niter=10, tol=1e-4):
    """optimization/cls_basic.py:82-141 (CG.setup/step/run)."""
    x = x0.copy()
    r = y - matvec(x)
    c = r.copy()
    kold = float(np.abs(r.dot(r.conj())))
    cost = [float(np.sqrt(kold))]
    iiter = 0
    while iiter < niter and kold > tol:
        Opc = matvec(c)
        cOpc = np.abs(c.dot(Opc.conj()))
        a = float(kold / cOpc)
        x += a * c
        r -= a * Opc
        k = float(np.abs(r.dot(r.conj())))
        b = float(k / kold)
        c = r + b * c
        kold = k
        iiter += 1
        cost.append(float(np.sqrt(kold)))
    return x, iiter, np.array(cost)


# --------------------------------------------------------------------------
# MPISecondDerivative  (basicoperators/SecondDerivative.py)  -- "next" row f2
# --------------------------------------------------------------------------
def second_derivative(x_flat: List[np.ndarray], dims: Tuple[int, ...], sampling: float = 1.0,
                      kind: str = "centered", edge: bool = False, adjoint: bool = False,
                      dtype=np.float64) -> List[np.ndarray]:
    """SecondDerivative.py:112-257 applied to a flat SCATTER vector (per-rank list)."""
    size = len(x_flat)
    dims = tuple(int(d) for d in (dims if np.ndim(dims) else (dims,)))
    shapes = local_shapes(dims, size, SCATTER, 0)
    x = reshaped_in(x_flat, shapes)
    N = dims[0]
    last = size - 1
    y = [np.zeros(s, dtype=dtype) for s in shapes]
    gh = lambda f, b: add_ghost_cells(x, 0, [f] * size if f is not None else None,   # noqa: E731
                                      [b] * size if b is not None else None)
    h2 = sampling ** 2
    if kind == "forward" and not adjoint:                  # :125-133
        g = gh(None, 2)
        for r in range(size):
            yf = g[r][2:] - 2 * g[r][1:-1] + g[r][:-2]
            if r == last:
                yf = np.append(yf, _z(min(N, 2), dims, yf.dtype), axis=0)
            y[r][:] = yf / h2
    elif kind == "forward" and adjoint:                    # :135-158
        g1 = gh(1, 1)
        g2 = gh(2, None)
        for r in range(size):
            if r == last:
                y[r][:-2] += x[r][:-2]
            else:
                y[r][:] += x[r][:]
            yf = g1[r][:-2]
            if r == 0:
                yf = np.append(_z(1, dims, yf.dtype), yf, axis=0)
            if r == last:
                yf = np.append(yf, _z(min(1, N - 1), dims, yf.dtype), axis=0)
            y[r][:] -= 2 * yf
            yf = g2[r][:-2]
            if r == 0:
                yf = np.append(_z(min(N, 2), dims, yf.dtype), yf, axis=0)
            y[r][:] += yf
            y[r][:] /= h2
    elif kind == "backward" and not adjoint:               # :160-168
        g = gh(2, None)
        for r in range(size):
            yb = g[r][2:] - 2 * g[r][1:-1] + g[r][:-2]
            if r == 0:
                yb = np.append(_z(min(N, 2), dims, yb.dtype), yb, axis=0)
            y[r][:] = yb / h2
    elif kind == "backward" and adjoint:                   # :170-191
        g1 = gh(None, 2)
        g2 = gh(1, 1)
        for r in range(size):
            yb = g1[r][2:]
            if r == last:
                yb = np.append(yb, _z(min(2, N), dims, yb.dtype), axis=0)
            y[r][:] += yb
            yb = 2 * g2[r][2:]
            if r == 0:
                yb = np.append(_z(1, dims, yb.dtype), yb, axis=0)
            if r == last:
                yb = np.append(yb, _z(min(1, N - 1), dims, yb.dtype), axis=0)
            y[r][:] -= yb
            if r == 0:
                y[r][2:] += x[r][2:]
            else:
                y[r][:] += x[r][:]
            y[r][:] /= h2
    elif kind == "centered" and not adjoint:               # :193-208
        g = gh(1, 1)
        for r in range(size):
            yc = g[r][2:] - 2 * g[r][1:-1] + g[r][:-2]
            if r == 0:
                yc = np.append(_z(1, dims, yc.dtype), yc, axis=0)
            if r == last:
                yc = np.append(yc, _z(min(1, N - 1), dims, yc.dtype), axis=0)
            y[r][:] = yc
            if edge:
                if r == 0:
                    y[r][0] = x[r][0] - 2 * x[r][1] + x[r][2]
                if r == last:
                    y[r][-1] = x[r][-3] - 2 * x[r][-2] + x[r][-1]
            y[r][:] /= h2
    elif kind == "centered" and adjoint:                   # :210-246
        g1 = gh(None, 2)
        g2 = gh(1, 1)
        g3 = gh(2, None)
        for r in range(size):
            yc = g1[r][1:-1]
            if r == last:
                yc = np.append(yc, _z(min(2, N), dims, yc.dtype), axis=0)
            y[r][:] += yc
            yc = 2 * g2[r][1:-1]
            if r == 0:
                yc = np.append(_z(1, dims, yc.dtype), yc, axis=0)
            if r == last:
                yc = np.append(yc, _z(min(1, N - 1), dims, yc.dtype), axis=0)
            y[r][:] -= yc
            yc = g3[r][1:-1]
            if r == 0:
                yc = np.append(_z(min(2, N), dims, yc.dtype), yc, axis=0)
            y[r][:] += yc
            if edge:
                if r == 0:
                    y[r][0] += x[r][0]
                    y[r][1] -= 2 * x[r][0]
                    y[r][2] += x[r][0]
                if r == last:
                    y[r][-3] += x[r][-1]
                    y[r][-2] -= 2 * x[r][-1]
                    y[r][-1] += x[r][-1]
            y[r][:] /= h2
    else:
        raise NotImplementedError("'kind' must be 'forward', 'centered' or 'backward'")
    return [a.ravel() for a in y]


def second_derivative_dense(N: int, sampling=1.0, kind="centered", edge=False) -> np.ndarray:
    """serial N x N second-derivative matrix along axis 0 (what the reference tests compare with)"""
    D = np.zeros((N, N))
    if kind == "forward":
        for i in range(N - 2):
            D[i, i], D[i, i + 1], D[i, i + 2] = 1, -2, 1
    elif kind == "backward":
        for i in range(2, N):
            D[i, i - 2], D[i, i - 1], D[i, i] = 1, -2, 1
    elif kind == "centered":
        for i in range(1, N - 1):
            D[i, i - 1], D[i, i], D[i, i + 1] = 1, -2, 1
        if edge:
            D[0, 0], D[0, 1], D[0, 2] = 1, -2, 1
            D[N - 1, N - 3], D[N - 1, N - 2], D[N - 1, N - 1] = 1, -2, 1
    else:
        raise NotImplementedError
    return D / sampling ** 2


def derivative_along_axis(x: np.ndarray, axis: int, D: np.ndarray) -> np.ndarray:
    """apply the dense stencil matrix D along `axis` of x (serial reference for Laplacian terms)"""
    return np.moveaxis(np.tensordot(D, x, axes=([1], [axis])), 0, axis)


# --------------------------------------------------------------------------
# MPIMDC (waveeqprocessing/MDC.py:12-74).  The FFT / Identity stages are third-party pylops operators that are
# not available here; their convention (orthonormal one-sided FFT with the positive frequencies scaled by
# sqrt(2), ifftshift of the time axis for two-sided data) is restated from pylops 2.x.  Pinned one level
# weaker than the hot path: the reference's own MPIMDC chain (MDC.py + MPIFredholm1 + MPILinearOperator
# algebra) is run unmodified over that restatement (tests/golden/refshim/pylops/signalprocessing) and this
# function reproduces it bit-exactly for complex128 (tests/test_golden.py::test_oracle_mdc); the third-party
# FFT convention itself stays unpinned.
# --------------------------------------------------------------------------
def _fft_real(x, nt, ifftshift_before):
    if ifftshift_before:
        x = np.fft.ifftshift(x, axes=0)
    y = np.fft.rfft(x, n=nt, axis=0, norm="ortho")
    y[1:1 + (nt - 1) // 2] *= np.sqrt(2)
    return y


def _fft_real_adj(y, nt, ifftshift_before):
    y = y.copy()
    y[1:1 + (nt - 1) // 2] /= np.sqrt(2)
    x = np.fft.irfft(y, n=nt, axis=0, norm="ortho")
    if ifftshift_before:
        x = np.fft.fftshift(x, axes=0)
    return x


def mdc(G_loc: List[np.ndarray], x: np.ndarray, nt: int, nv: int, twosided=True, adjoint=False,
        dt=1.0, dr=1.0, prescaled=False, conj=False) -> np.ndarray:
    """d = F1^H I1^H Fredholm1 I F m  (MDC.py:55-69) with G split over ranks along frequency; ``conj`` = the
    conjugated Fredholm operator of MDC.py:43-44 (matrix conj(G)).  Pinned against the reference's MPIMDC run over
    refshim's restated pylops FFT / Identity (tests/golden, key "mdc/")."""
    if conj:
        G_loc = [np.conj(g) for g in G_loc]
    nfmax = sum(g.shape[0] for g in G_loc)
    ns, nr = G_loc[0].shape[1:]
    nfft = int(np.ceil((nt + 1) / 2))
    sc = 1.0 if prescaled else dr * dt * np.sqrt(nt)
    Gs = [sc * g for g in G_loc]
    if not adjoint:
        X = _fft_real(np.real(x).reshape(nt, nr, nv), nt, twosided)[:nfmax]
        Y = fredholm1(Gs, X.ravel(), nv).reshape(nfmax, ns, nv)
        Yp = np.zeros((nfft, ns, nv), dtype=Y.dtype)
        Yp[:nfmax] = Y
        return _fft_real_adj(Yp, nt, False).ravel()
    Y = _fft_real(np.real(x).reshape(nt, ns, nv), nt, False)[:nfmax]
    X = fredholm1(Gs, Y.ravel(), nv, adjoint=True).reshape(nfmax, nr, nv)
    Xp = np.zeros((nfft, nr, nv), dtype=X.dtype)
    Xp[:nfmax] = X
    return _fft_real_adj(Xp, nt, twosided).ravel()


# ---------------------------------------------------------------------------------------------------------
# "next" rows f2/f3: MPIGradient (basicoperators/Gradient.py:88-119) and MPILaplacian
# (basicoperators/Laplacian.py:97-127) -- axis 0 is the distributed stencil above, the other axes are
# rank-local stencils (third-party pylops.FirstDerivative / SecondDerivative, restated as the dense
# operators of first_derivative_dense / second_derivative_dense applied along the axis of the rank's block).
# Pinned against the reference's own glue run over refshim/pylops/_derivatives.py (tests/golden).
# ---------------------------------------------------------------------------------------------------------
def _local_axis(x_flat, dims, axis, D_of_n, adjoint):
    """rank-local operator inside MPIBlockDiag: @reshaped(stacking=True) first re-partitions the flat
    vector to the operator's row blocks (BlockDiag.py:121-131, utils/decorators.py:70-85)"""
    shapes = local_shapes(tuple(dims), len(x_flat), SCATTER, 0)
    out = []
    for xl, shp in zip(reshaped_in(x_flat, shapes), shapes):
        D = D_of_n(shp[axis])
        out.append(derivative_along_axis(np.asarray(xl).reshape(shp), axis, D.T if adjoint else D).ravel())
    return out


def gradient(x_flat: List[np.ndarray], dims, sampling, kind="centered", edge=False) -> List[List[np.ndarray]]:
    """forward MPIGradient: list (one entry per axis) of per-rank flat arrays (Gradient.py:101-119)"""
    out = [first_derivative(x_flat, tuple(dims), sampling[0], kind, edge, 3, False)]
    for ax in range(1, len(dims)):
        out.append(_local_axis(x_flat, dims, ax, lambda n: first_derivative_dense(n, sampling[ax], kind, edge, 3), False))
    return out


def gradient_adjoint(y_flat: List[List[np.ndarray]], dims, sampling, kind="centered", edge=False) -> List[np.ndarray]:
    """adjoint MPIGradient: sum over axes of the per-axis adjoints (VStack.py:195-201)"""
    acc = first_derivative(y_flat[0], tuple(dims), sampling[0], kind, edge, 3, True)
    for ax in range(1, len(dims)):
        part = _local_axis(y_flat[ax], dims, ax, lambda n: first_derivative_dense(n, sampling[ax], kind, edge, 3), True)
        acc = [a + p for a, p in zip(acc, part)]
    return acc


def laplacian(x_flat: List[np.ndarray], dims, axes, weights, sampling, kind="centered", edge=False,
              adjoint=False) -> List[np.ndarray]:
    """MPILaplacian: weighted sum of second derivatives (Laplacian.py:97-127)"""
    acc = None
    for ax, w, s in zip(axes, weights, sampling):
        ax = ax % len(dims)
        if ax == 0:
            part = second_derivative(x_flat, tuple(dims), s, kind, edge, adjoint)
        else:
            part = _local_axis(x_flat, dims, ax, lambda n: second_derivative_dense(n, s, kind, edge), adjoint)
        part = [w * p for p in part]
        acc = part if acc is None else [a + p for a, p in zip(acc, part)]
    return acc


# ---------------------------------------------------------------------------------------------------------
# "next" row: ISTA / FISTA (optimization/cls_sparsity.py:140-343, 578-715) on a dense global operator.  The
# recurrences involve only global vectors and global norms, so the P-rank run equals the 1-rank run up to the
# reduction order.  Thresholds: third-party pylops.optimization.cls_sparsity (published formulas).  Pinned
# against the reference's own loops run over refshim (tests/golden, key "sparse/").
# ---------------------------------------------------------------------------------------------------------
def threshold(x: np.ndarray, thresh: float, kind: str) -> np.ndarray:
    a = np.abs(x)
    if kind == "soft":
        if np.iscomplexobj(x):
            return np.maximum(a - thresh, 0.0) * np.exp(1j * np.angle(x))
        return np.maximum(a - thresh, 0.0) * np.sign(x)
    if kind == "hard":
        return np.where(a <= np.sqrt(2 * thresh), 0, x)
    if kind == "half":
        arg = np.ones_like(x)
        nz = x != 0
        arg[nz] = (thresh / 8.0) * (a[nz] / 3.0) ** (-1.5)
        phi = 2.0 / 3.0 * np.arccos(np.clip(arg, -1, 1))
        x1 = 2.0 / 3.0 * x * (1 + np.cos(2.0 * np.pi / 3.0 - phi))
        return np.where(a <= (54 ** (1.0 / 3.0) / 4.0) * thresh ** (2.0 / 3.0), 0, x1)
    raise ValueError(f"threshkind must be hard, soft, half, got {kind}")


def ista(A: np.ndarray, y: np.ndarray, x0: np.ndarray, niter: int, eps: float, alpha: float, tol: float = 1e-10,
         threshkind: str = "soft", fista: bool = False):
    """returns (x, iiter, cost); ``fista=True`` adds the momentum of cls_sparsity.py:636-644"""
    thresh = eps * alpha * 0.5                                            # :246
    x = x0.copy()
    z = x.copy()
    t = 1.0
    cost, iiter, xupdate = [], 0, np.inf
    while iiter < niter and xupdate > tol:                                # :376 / :693
        xold = x.copy()
        res = y - A @ (z if fista else x)
        x = threshold((z if fista else x) + alpha * (A.conj().T @ res), thresh, threshkind)
        if fista:
            told = t
            t = (1.0 + np.sqrt(1.0 + 4.0 * t ** 2)) / 2.0
            z = x + ((told - 1.0) / t) * (x - xold)
            res = y - A @ x                                               # :652 cost on the new x
        xupdate = np.linalg.norm(x - xold)
        cost.append(0.5 * np.linalg.norm(res) ** 2 + eps * np.sum(np.abs(x)))
        iiter += 1
    return x, iiter, np.array(cost)


def power_iteration(A: np.ndarray, niter: int = 10, tol: float = 1e-5, seed: int = 0):
    """optimization/eigs.py:10-98 (start vector differs: the reference draws it from the global NumPy RNG)"""
    rng = np.random.default_rng(seed)
    b = rng.random(A.shape[1]).astype(A.dtype)
    b /= np.linalg.norm(b)
    old = 0.0
    for it in range(niter):
        b1 = A @ b
        eig = np.vdot(b, b1)
        b = b1 / np.linalg.norm(b1)
        if np.abs(eig - old) < tol * eig:
            break
        old = eig
    return eig, b, it + 1

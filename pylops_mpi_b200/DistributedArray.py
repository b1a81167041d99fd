"""B200-native ``DistributedArray`` with the reference's API surface
(pylops_mpi/DistributedArray.py:26-959).

Storage is a device buffer (a torch CUDA tensor -- plumbing only); every
arithmetic / reduction method launches a hand-written sm_100a kernel from
libb200lops through the C ABI, every collective is NCCL over NVLink.  Partition
metadata is computed once at construction (closed formulas), so the steady
state issues zero host collectives -- the reference issues an allgather or
allreduce of shapes for every temporary it creates (:345-358, :523-539).
"""
from __future__ import annotations

from enum import Enum
from numbers import Integral
from typing import List, Optional, Sequence, Tuple, Union

import math

import numpy as np
import torch

from . import _lib
from .comm import Comm, COMM_WORLD, resolve, SUM, MAX, MIN
from .Distributed import (DistributedMixIn, allreduce_, allgatherv, bcast_, group, send, recv)
from .utils.partition import local_split_sizes, offsets, repartition_plan

__all__ = ["Partition", "local_split", "subcomm_split", "DistributedArray"]


class Partition(Enum):
    """DistributedArray.py:26-39"""
    BROADCAST = "Broadcast"
    UNSAFE_BROADCAST = "UnsafeBroadcast"
    SCATTER = "Scatter"


_BCAST = (Partition.BROADCAST, Partition.UNSAFE_BROADCAST)


def _tup(v) -> Tuple[int, ...]:
    if isinstance(v, (Integral, np.integer)):
        return (int(v),)
    return tuple(int(i) for i in v)


def local_split(global_shape: Tuple, base_comm, partition: Partition, axis: int) -> Tuple:
    """DistributedArray.py:42-71 (integer bookkeeping, bit-exact)."""
    base_comm = resolve(base_comm)
    if partition in _BCAST:
        return tuple(global_shape)
    local_shape = list(global_shape)
    rank, size = base_comm.Get_rank(), base_comm.Get_size()
    if rank < (global_shape[axis] % size):
        local_shape[axis] = global_shape[axis] // size + 1
    else:
        local_shape[axis] = global_shape[axis] // size
    return tuple(local_shape)


_KERNEL_DTYPES = (torch.float32, torch.float64, torch.complex64, torch.complex128)


def subcomm_split(mask, comm=COMM_WORLD):
    """DistributedArray.py:74-100 (cached: one Split per distinct mask)."""
    return resolve(comm).split_by_mask(mask)


def _as_device(value, dtype: torch.dtype, device) -> torch.Tensor:
    if isinstance(value, DistributedArray):
        value = value.local_array
    if isinstance(value, torch.Tensor):
        return value.to(device=device, dtype=dtype, non_blocking=True)
    return torch.as_tensor(np.asarray(value), device=device).to(dtype)


class DistributedArray(DistributedMixIn):
    """See pylops_mpi/DistributedArray.py:103-182 for the parameters.  ``engine``
    is accepted and ignored (there is one engine: B200); ``base_comm_nccl`` is
    accepted and ignored (NCCL lives inside ``base_comm``)."""

    def __init__(self, global_shape: Union[Tuple, Integral],
                 base_comm: Optional[Comm] = COMM_WORLD,
                 base_comm_nccl=None,
                 partition: Partition = Partition.SCATTER, axis: int = 0,
                 local_shapes: Optional[List[Union[Tuple, Integral]]] = None,
                 mask: Optional[List[Integral]] = None,
                 engine: Optional[str] = "b200",
                 dtype=np.float64,
                 _buffer: Optional[torch.Tensor] = None, _trusted: bool = False):
        global_shape = _tup(global_shape)
        if len(global_shape) <= axis:
            raise IndexError(f"Axis {axis} out of range for DistributedArray "
                             f"of shape {global_shape}")
        if not _trusted and partition not in Partition:
            raise ValueError(f"Should be either {Partition.BROADCAST}, "
                             f"{Partition.UNSAFE_BROADCAST} or {Partition.SCATTER}")
        self._tdtype = _lib.torch_dtype(dtype)
        self.dtype = _lib.numpy_dtype(dtype)
        self._global_shape = global_shape
        self._base_comm = resolve(base_comm)
        self._base_comm_nccl = None
        self._partition = partition
        self._axis = axis
        self._mask = mask
        self._sub_comm = self._base_comm if mask is None else subcomm_split(mask, self._base_comm)
        size, rank = self._base_comm.Get_size(), self._base_comm.Get_rank()
        if local_shapes is not None:
            if not (_trusted and type(local_shapes[0]) is tuple):     # operator temporaries pass validated tuples
                local_shapes = [_tup(s) for s in local_shapes]
            if not _trusted:  # internal constructions pass shapes derived from validated arrays
                self._check_local_shapes(local_shapes)
            self._local_shapes = local_shapes
        elif partition in _BCAST:
            self._local_shapes = [global_shape] * size
        else:
            ext = local_split_sizes(global_shape[axis], size)
            self._local_shapes = [global_shape[:axis] + (e,) + global_shape[axis + 1:] for e in ext]
        self._local_shape = self._local_shapes[rank]
        self._engine = "b200"
        if _buffer is not None:
            if tuple(_buffer.shape) != self._local_shape or _buffer.dtype != self._tdtype:
                raise ValueError("internal buffer does not match the local shape / dtype")
            self._local_array = _buffer
        else:
            _lib.ctx()  # fail loudly without a CUDA device / the extension
            self._local_array = torch.empty(self._local_shape, dtype=self._tdtype, device="cuda")

    @classmethod
    def _internal(cls, global_shape, local_shapes, base_comm, tdtype, buffer=None, partition=Partition.SCATTER,
                  axis=0, mask=None):
        """operator temporaries: every argument is already validated and normalised (tuples of ints, resolved
        communicator, torch dtype) -- skips the checks of ``__init__`` (a few microseconds each, three per apply)"""
        self = object.__new__(cls)
        self._tdtype = tdtype
        self.dtype = _lib.numpy_dtype(tdtype)
        self._global_shape = global_shape
        self._base_comm = base_comm
        self._base_comm_nccl = None
        self._partition = partition
        self._axis = axis
        self._mask = mask
        self._sub_comm = base_comm if mask is None else subcomm_split(mask, base_comm)
        self._local_shapes = local_shapes
        self._local_shape = local_shapes[base_comm.Get_rank()]
        self._engine = "b200"
        if buffer is None:
            _lib.ctx()  # fail loudly without a CUDA device / the extension
            buffer = torch.empty(self._local_shape, dtype=tdtype, device="cuda")
        self._local_array = buffer
        return self

    # ---- element access ------------------------------------------------------
    def __getitem__(self, index):
        return self._local_array[index]

    def __setitem__(self, index, value):
        """DistributedArray.py:187-222: BROADCAST writes are re-broadcast from rank 0."""
        if isinstance(value, DistributedArray):
            value = value.local_array
        if self.partition is Partition.BROADCAST and self.size > 1:
            view = self._local_array[index]
            buf = torch.empty(view.shape, dtype=self._tdtype, device=self._local_array.device)
            if self.rank == 0:
                buf[...] = value if not isinstance(value, np.ndarray) else torch.as_tensor(value)
            bcast_(self._base_comm, buf, root=0)
            self._local_array[index] = buf
        else:
            if isinstance(value, np.ndarray):
                value = torch.as_tensor(value)
            self._local_array[index] = value

    # ---- properties (DistributedArray.py:224-368) ----------------------------------
    @property
    def global_shape(self):
        return self._global_shape

    @property
    def base_comm(self):
        return self._base_comm

    @property
    def base_comm_nccl(self):
        return self._base_comm_nccl

    @property
    def local_shape(self):
        return self._local_shape

    @property
    def mask(self):
        return self._mask

    @property
    def engine(self):
        return self._engine

    @property
    def local_array(self):
        return self._local_array

    @property
    def rank(self):
        return self._base_comm.Get_rank()

    @property
    def size(self):
        return self._base_comm.Get_size()

    @property
    def axis(self):
        return self._axis

    @property
    def ndim(self):
        return len(self._global_shape)

    @property
    def partition(self):
        return self._partition

    @property
    def local_shapes(self):
        """cached at construction (the reference allgathers on every access, :345-358)"""
        return list(self._local_shapes)

    @property
    def sub_comm(self):
        return self._sub_comm

    # ---- gather / scatter -----------------------------------------------------------
    def asarray(self, masked: bool = False) -> torch.Tensor:
        """Global view gathered on every rank (DistributedArray.py:370-405); returns a
        device tensor (``.cpu().numpy()`` for a host array)."""
        if self.partition in _BCAST:
            return self._local_array
        comm = self._sub_comm if masked else self._base_comm
        if comm.size == 1:
            return self._local_array
        if masked:
            members = [r for r in range(self.size) if self._mask[r] == self._mask[self.rank]]
            shapes = [self._local_shapes[r] for r in members]
        else:
            shapes = self._local_shapes
        counts = [int(np.prod(s)) for s in shapes]
        flat = allgatherv(comm, self._local_array.contiguous().view(-1), counts)
        if self._axis == 0:
            tot = sum(s[0] for s in shapes)
            return flat.view((tot,) + tuple(shapes[0][1:]))
        parts, off = [], 0
        for s, c in zip(shapes, counts):
            parts.append(flat[off:off + c].view(s))
            off += c
        return torch.cat(parts, dim=self._axis)

    @classmethod
    def to_dist(cls, x, base_comm=COMM_WORLD, base_comm_nccl=None,
                partition: Partition = Partition.SCATTER, axis: int = 0,
                local_shapes: Optional[List[Tuple]] = None,
                mask: Optional[List[Integral]] = None) -> "DistributedArray":
        """The reference's "Scatter": every rank slices its own block of the replicated
        global array (DistributedArray.py:407-460).  ``x`` may be a NumPy array, a host
        tensor or a device tensor."""
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x))
        dist_array = cls(global_shape=tuple(x.shape), base_comm=base_comm, partition=partition,
                         axis=axis, local_shapes=local_shapes, mask=mask, dtype=x.dtype)
        if partition in _BCAST:
            dist_array[:] = x
        else:
            ext = offsets([s[axis] for s in dist_array._local_shapes])
            sl = [slice(None)] * x.dim()
            sl[axis] = slice(ext[dist_array.rank], ext[dist_array.rank + 1])
            dist_array._local_array.copy_(x[tuple(sl)], non_blocking=True)
        return dist_array

    def redistribute(self, axis: int) -> "DistributedArray":
        """All-to-all re-partition along another axis (DistributedArray.py:462-521)."""
        if self.axis == axis or self.partition is not Partition.SCATTER:
            return self
        out = DistributedArray(global_shape=self.global_shape, base_comm=self.base_comm,
                               mask=self.mask, axis=axis, dtype=self.dtype)
        counts_from = [s[self.axis] for s in self._local_shapes]
        counts_to = [s[axis] for s in out._local_shapes]
        off_to = offsets(counts_to)
        send_bufs, recv_bufs = [], []
        for r in range(self.size):
            sl = [slice(None)] * self.ndim
            sl[axis] = slice(off_to[r], off_to[r + 1])
            send_bufs.append(self._local_array[tuple(sl)].contiguous())
            shp = list(self.global_shape)
            shp[self.axis] = counts_from[r]
            shp[axis] = counts_to[self.rank]
            recv_bufs.append(torch.empty(shp, dtype=self._tdtype, device=self._local_array.device))
        if self.size > 1:
            with group(self._base_comm):
                for r in range(self.size):
                    if r != self.rank:
                        send(self._base_comm, send_bufs[r], r)
                        recv(self._base_comm, recv_bufs[r], r)
        recv_bufs[self.rank] = send_bufs[self.rank]
        out._local_array.copy_(torch.cat(recv_bufs, dim=self.axis))
        return out

    # ---- checks (DistributedArray.py:523-554) ------------------------------------------
    def _check_local_shapes(self, local_shapes):
        if local_shapes:
            size, rank = self._base_comm.Get_size(), self._base_comm.Get_rank()
            if len(local_shapes) != size:
                raise ValueError(f"Length of local shapes is not equal to number of processes; "
                                 f"{len(local_shapes)} != {size}")
            if self._partition in _BCAST and local_shapes[rank] != self._global_shape:
                raise ValueError(f"Local shape is not equal to global shape at rank = {rank};"
                                 f"{local_shapes[rank]} != {self._global_shape}")
            elif self._partition is Partition.SCATTER:
                local_shape = local_shapes[rank]
                # the full list is known on every rank: the Allreduce of :536 is a local sum
                if sum(s[self._axis] for s in local_shapes) != self._global_shape[self._axis] or \
                        not np.array_equal(np.delete(local_shape, self._axis),
                                           np.delete(self._global_shape, self._axis)):
                    raise ValueError(f"Local shapes don't align with the global shape;"
                                     f"{local_shapes} != {self._global_shape}")

    def _check_partition_shape(self, dist_array):
        if self.partition != dist_array.partition:
            raise ValueError("Partition of both the arrays must be same")
        if self.local_shape != dist_array.local_shape:
            raise ValueError(f"Local Array Shape Mismatch - "
                             f"{self.local_shape} != {dist_array.local_shape}")

    def _check_mask(self, dist_array):
        if not np.array_equal(self.mask, dist_array.mask):
            raise ValueError("Mask of both the arrays must be same")

    # ---- construction helpers -------------------------------------------------------
    def _like(self, buffer: Optional[torch.Tensor] = None, mask="same", dtype=None) -> "DistributedArray":
        return DistributedArray(global_shape=self._global_shape, base_comm=self._base_comm,
                                partition=self._partition, axis=self._axis,
                                local_shapes=self._local_shapes,
                                mask=self._mask if mask == "same" else mask,
                                dtype=self._tdtype if dtype is None else dtype, _buffer=buffer, _trusted=True)

    def _cont(self) -> torch.Tensor:
        a = self._local_array
        return a if a.is_contiguous() else a.contiguous()

    def _pair(self, other: "DistributedArray") -> Tuple[torch.Tensor, torch.Tensor]:
        """contiguous local buffers of ``self`` and ``other`` in their COMMON dtype (the NumPy promotion of the
        reference's ``self.local_array + other.local_array``): the single-dtype kernels never see mixed buffers"""
        x, y = self._cont(), other._cont()
        if x.dtype != y.dtype:
            dt = torch.promote_types(x.dtype, y.dtype)
            x, y = x.to(dt), y.to(dt)
        return x, y

    def _mine_from(self, out: torch.Tensor) -> "DistributedArray":
        """wrap a result computed in the promoted dtype as an array of THIS array's dtype: the reference allocates
        every binary-op result with ``dtype=self.dtype`` and assigns the NumPy-promoted expression into it
        (DistributedArray.py:603-652), i.e. float32 + float64 -> float32, real + complex -> real (+ ComplexWarning)"""
        if out.dtype != self._tdtype:
            if out.dtype.is_complex and not self._tdtype.is_complex:
                import warnings
                warnings.warn("Casting complex values to real discards the imaginary part",
                              np.exceptions.ComplexWarning, stacklevel=3)
                out = out.real
            out = out.to(self._tdtype).contiguous()
        return self._like(out)

    def _as_mine(self, other: "DistributedArray") -> torch.Tensor:
        """``other``'s contiguous buffer cast to THIS array's dtype.  In-place updates keep the left operand's
        dtype: the reference writes them as ``self[:] = self.local_array + other.local_array``
        (DistributedArray.py:591-592), i.e. a NumPy ``__setitem__`` cast -- complex into real DISCARDS the imaginary
        part with a ComplexWarning (this is what lets a real model be updated with the complex-typed, zero-imaginary
        output of ``MDCop.H`` in CGLS, tutorials/mdd.py:190-194)."""
        y = other._cont()
        if y.dtype != self._tdtype:
            if y.dtype.is_complex and not self._tdtype.is_complex:
                import warnings
                warnings.warn("Casting complex values to real discards the imaginary part",
                              np.exceptions.ComplexWarning, stacklevel=3)
                y = y.real
            y = y.to(self._tdtype).contiguous()
        return y

    def _lincomb(self, a, x: torch.Tensor, b=None, y: Optional[torch.Tensor] = None,
                 out: Optional[torch.Tensor] = None, conj_x: bool = False) -> torch.Tensor:
        """out = a*op(x) + b*y on the device (b2_lincomb)"""
        if out is None:
            out = torch.empty_like(x)
        if (y is not None and y.dtype != x.dtype) or out.dtype != x.dtype:
            raise TypeError(f"b2_lincomb needs one dtype, got {x.dtype} / {None if y is None else y.dtype} / {out.dtype}")
        n = x.numel()
        if n:
            _lib.check(_lib.lib.b2_lincomb(_lib.ctx(), out.data_ptr(), _lib.cpair(a), x.data_ptr(),
                                           _lib.cpair(b) if y is not None else None,
                                           y.data_ptr() if y is not None else None, n,
                                           _lib.code(x.dtype), int(conj_x), _lib.stream()),
                       "b2_lincomb")
        return out

    # ---- arithmetic (DistributedArray.py:574-652) ----------------------------------------
    def __neg__(self):
        return self._like(self._lincomb(-1.0, self._cont()))

    def __add__(self, x):
        return self.add(x)

    def __iadd__(self, x):
        return self.iadd(x)

    def __sub__(self, x):
        self._check_partition_shape(x)
        self._check_mask(x)
        a, b = self._pair(x)
        return self._mine_from(self._lincomb(1.0, a, -1.0, b))

    def _iupdate(self, other: "DistributedArray", sign: float):
        """self <- self + sign * other.  Same dtype: one in-place kernel.  Mixed dtypes: the reference evaluates
        ``self[:] = self.local_array + other.local_array`` (DistributedArray.py:619-624): the sum in the promoted
        dtype, then a NumPy __setitem__ cast into self -- reproduced exactly (sum first, cast second)"""
        if other._tdtype == self._tdtype:
            a = self._cont()
            self._lincomb(1.0, a, sign, other._cont(), out=a)
            if a is not self._local_array:
                self._local_array.copy_(a)
            return self
        xa, xb = self._pair(other)
        res = self._mine_from(self._lincomb(1.0, xa, sign, xb))
        self._local_array.copy_(res._local_array)
        return self

    def __isub__(self, x):
        self._check_partition_shape(x)
        self._check_mask(x)
        return self._iupdate(x, -1.0)

    def __mul__(self, x):
        return self.multiply(x)

    def __rmul__(self, x):
        return self.multiply(x)

    def add(self, dist_array):
        self._check_partition_shape(dist_array)
        self._check_mask(dist_array)
        a, b = self._pair(dist_array)
        return self._mine_from(self._lincomb(1.0, a, 1.0, b))

    def iadd(self, dist_array):
        self._check_partition_shape(dist_array)
        self._check_mask(dist_array)
        return self._iupdate(dist_array, 1.0)

    def multiply(self, dist_array):
        if isinstance(dist_array, DistributedArray):
            self._check_partition_shape(dist_array)
            self._check_mask(dist_array)
            x, y = self._pair(dist_array)
            out = torch.empty_like(x)
            if x.numel():
                _lib.check(_lib.lib.b2_mul(_lib.ctx(), out.data_ptr(), x.data_ptr(), y.data_ptr(),
                                           x.numel(), _lib.code(x.dtype), 0, _lib.stream()), "b2_mul")
            return self._mine_from(out)
        scalar = complex(dist_array)
        x = self._cont()
        if scalar.imag != 0.0 and not self._tdtype.is_complex:
            # the product is formed in the matching complex dtype, then assigned into a self.dtype array
            x = x.to(torch.complex64 if self._tdtype in (torch.float32, torch.bfloat16) else torch.complex128)
        return self._mine_from(self._lincomb(scalar, x))

    # fused updates used by the solvers (no temporaries): self <- self + a*x ; self <- x + b*self
    def axpy_(self, a, x: "DistributedArray"):
        s = self._cont()
        self._lincomb(a, self._as_mine(x), 1.0, s, out=s)
        if s is not self._local_array:
            self._local_array.copy_(s)
        return self

    def xpby_(self, x: "DistributedArray", b):
        s = self._cont()
        self._lincomb(1.0, self._as_mine(x), b, s, out=s)
        if s is not self._local_array:
            self._local_array.copy_(s)
        return self

    def scale_(self, a):
        s = self._cont()
        self._lincomb(a, s, out=s)
        if s is not self._local_array:
            self._local_array.copy_(s)
        return self

    # ---- reductions (DistributedArray.py:654-807) ----------------------------------------
    def _scatter_view(self) -> torch.Tensor:
        """flat local data with each global element counted once: BROADCAST arrays are
        re-scattered (to_dist on the flattened... first axis) as :678-681 / :792-793 do"""
        a = self._cont()
        if self.partition in _BCAST:
            ext = offsets(local_split_sizes(a.shape[0], self.size))
            a = a[ext[self.rank]:ext[self.rank + 1]]
        return a.reshape(-1)

    def _dot_device(self, dist_array, vdot: bool = False) -> torch.Tensor:
        """device float64[2] = (re, im) of the global dot product (no host sync)"""
        x, y = self._scatter_view(), dist_array._scatter_view()
        if x.dtype != y.dtype:                     # NumPy promotes; the kernel takes ONE dtype code
            dt = torch.promote_types(x.dtype, y.dtype)
            x, y = x.to(dt), y.to(dt)
        out = torch.empty(2, dtype=torch.float64, device=x.device)
        _lib.check(_lib.lib.b2_dot(_lib.ctx(), x.data_ptr() if x.numel() else None,
                                   y.data_ptr() if y.numel() else None, x.numel(),
                                   _lib.code(x.dtype), int(vdot), out.data_ptr(), _lib.stream()),
                   "b2_dot")
        return allreduce_(self._sub_comm, out, SUM)

    def dot(self, dist_array, vdot: bool = False):
        """Distributed dot product (DistributedArray.py:654-686).  Returns a 1-element
        host NumPy array of this array's dtype, like the reference's MPI path
        (utils/_mpi.py:102-106); accumulation is float64 on the device."""
        self._check_partition_shape(dist_array)
        self._check_mask(dist_array)
        res = self._dot_device(dist_array, vdot).cpu().numpy()
        if self._tdtype.is_complex:
            return np.array([complex(res[0], res[1])], dtype=self.dtype)
        if self._tdtype is torch.bfloat16:
            return np.array([res[0]], dtype=np.float32)
        return np.array([res[0]], dtype=self.dtype)

    def _norm_device(self, ord=None) -> Tuple[torch.Tensor, float]:
        ord = 2 if ord is None else ord
        if ord in ("fro", "nuc"):
            raise ValueError(f"norm-{ord} not possible for vectors")
        x = self._scatter_view()
        if ord == 0:
            kind, op, p, root = _lib.NRM_COUNT_NONZERO, SUM, 0.0, 1.0
        elif ord == np.inf:
            kind, op, p, root = _lib.NRM_MAX_ABS, MAX, 0.0, 1.0
        elif ord == -np.inf:
            kind, op, p, root = _lib.NRM_MIN_ABS, MIN, 0.0, 1.0
        elif ord == 1:
            kind, op, p, root = _lib.NRM_SUM_ABS, SUM, 1.0, 1.0
        elif ord == 2:
            kind, op, p, root = _lib.NRM_SUM_SQ, SUM, 2.0, 0.5
        else:
            kind, op, p, root = _lib.NRM_SUM_POW, SUM, float(ord), 1.0 / float(ord)
        out = torch.empty(1, dtype=torch.float64, device=x.device)
        _lib.check(_lib.lib.b2_norm_partial(_lib.ctx(), x.data_ptr() if x.numel() else None, x.numel(),
                                            _lib.code(x.dtype), kind, p, out.data_ptr(), _lib.stream()),
                   "b2_norm_partial")
        return allreduce_(self._sub_comm, out, op), root

    @staticmethod
    def _norm_kind(ord):
        ord = 2 if ord is None else ord
        if ord in ("fro", "nuc"):
            raise ValueError(f"norm-{ord} not possible for vectors")
        if ord == 0:
            return _lib.NRM_COUNT_NONZERO, SUM, 0.0, 1.0
        if ord == np.inf:
            return _lib.NRM_MAX_ABS, MAX, 0.0, 1.0
        if ord == -np.inf:
            return _lib.NRM_MIN_ABS, MIN, 0.0, 1.0
        if ord == 1:
            return _lib.NRM_SUM_ABS, SUM, 1.0, 1.0
        if ord == 2:
            return _lib.NRM_SUM_SQ, SUM, 2.0, 0.5
        return _lib.NRM_SUM_POW, SUM, float(ord), 1.0 / float(ord)

    @staticmethod
    def _axis_partials(a: torch.Tensor, axis: int, kind: int, p: float) -> torch.Tensor:
        """float64 partials of the norm along ``axis`` of the local block (b2_norm_axis; no eager torch pass)"""
        a = a if a.is_contiguous() else a.contiguous()
        shp = tuple(a.shape)
        n_outer = int(np.prod(shp[:axis])) if axis else 1
        n_axis = int(shp[axis])
        n_inner = int(np.prod(shp[axis + 1:])) if axis + 1 < len(shp) else 1
        out = torch.empty(shp[:axis] + shp[axis + 1:], dtype=torch.float64, device=a.device)
        if out.numel():
            if n_axis == 0:
                out.fill_(float("inf") if kind == _lib.NRM_MIN_ABS else 0.0)
            else:
                _lib.check(_lib.lib.b2_norm_axis(_lib.ctx(), a.data_ptr(), n_outer, n_axis, n_inner, _lib.code(a.dtype),
                                                 kind, p, out.data_ptr(), _lib.stream()), "b2_norm_axis")
        return out

    def _compute_vector_norm(self, local_array, axis: int, ord=None):
        """axis-wise variant (DistributedArray.py:688-758) for arrays partitioned along ``axis``: per-column
        float64 partials on the device (b2_norm_axis), one Allreduce of the partial vector, root on the device"""
        kind, op, p, root = self._norm_kind(ord)
        part = self._axis_partials(local_array, axis, kind, p)
        red = allreduce_(self._sub_comm, part.contiguous(), op)
        return red.pow(root) if root != 1.0 else red

    def norm(self, ord=None, axis: Optional[int] = None):
        """Distributed vector norm (DistributedArray.py:774-807); float64 result
        (the reference's float_power promotion, :755).  axis=None returns a
        1-element host array."""
        if axis is None:
            val, root = self._norm_device(ord)
            res = val.cpu().numpy()
            return np.power(res, root) if root != 1.0 else res
        if axis >= self.ndim:
            raise ValueError(f"axis={axis} is out of range for array of dimension {self.ndim}")
        if self.partition in _BCAST:
            x = DistributedArray.to_dist(self._local_array, base_comm=self._base_comm)
        else:
            x = self
        if self.axis != axis:
            norm_axis = self.axis - 1 if axis < self.axis else self.axis
            kind, _op, p, root = self._norm_kind(ord)
            loc = self._axis_partials(x.local_array, axis, kind, p)       # the whole axis is local: take the root here
            if root != 1.0:
                loc = loc.pow(root)
            parts = self._allgather(self._base_comm, None, loc.contiguous())
            return torch.cat(parts, dim=norm_axis)
        return x._compute_vector_norm(x.local_array, axis=axis, ord=ord)

    # ---- copies (DistributedArray.py:760-874) ----------------------------------------------
    def zeros_like(self):
        """note: like the reference (:763-770) the mask is NOT propagated"""
        arr = self._like(mask=None)
        n = arr._local_array.numel()
        if self._tdtype not in _KERNEL_DTYPES:      # integer / bf16 arrays: plain memset
            arr._local_array.zero_()
        elif n:
            _lib.check(_lib.lib.b2_fill(_lib.ctx(), arr._local_array.data_ptr(), _lib.cpair(0.0), n,
                                        _lib.code(self._tdtype), _lib.stream()), "b2_fill")
        return arr

    def conj(self):
        if not self._tdtype.is_complex:
            return self.copy()
        return self._like(self._lincomb(1.0, self._cont(), conj_x=True))

    def copy(self):
        if self._tdtype not in _KERNEL_DTYPES:
            return self._like(self._local_array.clone(memory_format=torch.contiguous_format))
        return self._like(self._lincomb(1.0, self._cont()))      # one b2_lincomb pass (out = 1 * x)

    def ravel(self, order: Optional[str] = "C"):
        if order != "C":
            raise NotImplementedError("only C-order ravel is supported on device")
        local_shapes = [(int(np.prod(s)),) for s in self._local_shapes]
        return DistributedArray(global_shape=int(np.prod(self._global_shape)),
                                base_comm=self._base_comm, local_shapes=local_shapes,
                                mask=self._mask, partition=self._partition, dtype=self._tdtype,
                                _buffer=self._cont().reshape(-1).clone(), _trusted=True)

    def _ravel_view(self):
        """flattened DistributedArray SHARING this array's buffer (internal: used on
        operator temporaries where the reference's ravel() copy is pure overhead)"""
        local_shapes = [(math.prod(s),) for s in self._local_shapes]
        return DistributedArray._internal((math.prod(self._global_shape),), local_shapes, self._base_comm,
                                          self._tdtype, buffer=self._cont().reshape(-1), partition=self._partition,
                                          mask=self._mask)

    def empty_like(self):
        return self._like()

    # ---- halo exchange (DistributedArray.py:876-953) ---------------------------------------
    def add_ghost_cells(self, cells_front: Optional[int] = None, cells_back: Optional[int] = None):
        """Returns the local array extended with ``cells_front`` cells of rank-1 and
        ``cells_back`` cells of rank+1 along the partition axis."""
        ax = self._axis
        rank, size = self.rank, self.size
        loc = self._local_array
        front = back = None
        if cells_front is not None:
            total = self._base_comm.allgather(int(cells_front)) + [0]
            to_next = total[rank + 1]
            if rank != size - 1 and to_next != 0 and to_next > self.local_shape[ax]:
                raise ValueError(f"Local Shape at rank={rank} along axis={ax} "
                                 f"should be > {to_next}: dim({ax}) "
                                 f"{self.local_shape[ax]} < {to_next}; "
                                 f"to achieve this use NUM_PROCESSES <= "
                                 f"{max(1, self.global_shape[ax] // to_next)}")
            snd = loc.narrow(ax, loc.shape[ax] - to_next, to_next).contiguous() if to_next else None
            if rank != 0 and total[rank] != 0:
                shp = list(self._local_shapes[rank - 1])
                shp[ax] = total[rank]
                front = torch.empty(shp, dtype=self._tdtype, device=loc.device)
            if size > 1:
                with group(self._base_comm):
                    if rank != size - 1 and snd is not None:
                        send(self._base_comm, snd, rank + 1)
                    if front is not None:
                        recv(self._base_comm, front, rank - 1)
        if cells_back is not None:
            total = self._base_comm.allgather(int(cells_back)) + [0]
            to_prev = total[rank - 1] if rank != 0 else 0
            if rank != 0 and to_prev != 0 and to_prev > self.local_shape[ax]:
                raise ValueError(f"Local Shape at rank={rank} along axis={ax} "
                                 f"should be > {to_prev}: dim({ax}) "
                                 f"{self.local_shape[ax]} < {to_prev}; "
                                 f"to achieve this use NUM_PROCESSES <= "
                                 f"{max(1, self.global_shape[ax] // to_prev)}")
            snd = loc.narrow(ax, 0, to_prev).contiguous() if to_prev else None
            if rank != size - 1 and total[rank] != 0:
                shp = list(self._local_shapes[rank + 1])
                shp[ax] = total[rank]
                back = torch.empty(shp, dtype=self._tdtype, device=loc.device)
            if size > 1:
                with group(self._base_comm):
                    if rank != 0 and snd is not None:
                        send(self._base_comm, snd, rank - 1)
                    if back is not None:
                        recv(self._base_comm, back, rank + 1)
        parts = [p for p in (front, loc, back) if p is not None]
        return torch.cat(parts, dim=ax) if len(parts) > 1 else loc.clone()

    # ---- flat re-partition used by the operators (replaces @reshaped's ghost shuffle) ------
    def _repartition_flat(self, dst_sizes: Sequence[int]) -> torch.Tensor:
        """Move this flat SCATTER vector to the 1-D partition ``dst_sizes``; returns the
        new flat local buffer.  No-op (a view) when the partitions already agree --
        the common case for row-aligned shapes."""
        src_sizes = [math.prod(s) for s in self._local_shapes]
        flat = self._cont().reshape(-1)
        if src_sizes == dst_sizes:
            return flat
        dst_sizes = [int(d) for d in dst_sizes]
        sends, recvs = repartition_plan(src_sizes, dst_sizes, self.rank)
        out = torch.empty(dst_sizes[self.rank], dtype=self._tdtype, device=flat.device)
        with group(self._base_comm):
            for peer, off, cnt in sends:
                if peer != self.rank:
                    send(self._base_comm, flat[off:off + cnt], peer)
            for peer, off, cnt in recvs:
                if peer != self.rank:
                    recv(self._base_comm, out[off:off + cnt], peer)
        for (peer, soff, cnt) in sends:
            if peer == self.rank:
                doff = [r for r in recvs if r[0] == self.rank][0][1]
                out[doff:doff + cnt].copy_(flat[soff:soff + cnt])
        return out

    def __repr__(self):
        return f"<DistributedArray with global shape={self.global_shape}, " \
               f"local shape={self.local_shape}" \
               f", dtype={self.dtype}, " \
               f"processes={[i for i in range(self.size)]})> "


def __getattr__(name):
    # import-path parity: the reference defines StackedDistributedArray in this module (DistributedArray.py:962);
    # here it lives in StackedArray.py, which imports this module -> resolve lazily
    if name == "StackedDistributedArray":
        from .StackedArray import StackedDistributedArray
        return StackedDistributedArray
    raise AttributeError(f"module {__name__!r} has no attribute {name!r}")

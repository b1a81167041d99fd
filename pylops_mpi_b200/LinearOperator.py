"""``MPILinearOperator`` and its lazy algebra with the reference's semantics
(pylops_mpi/LinearOperator.py:14-438).  Pure host glue: subclasses implement
``_matvec`` / ``_rmatvec`` on :class:`DistributedArray`; scaling / conjugation
of results runs in libb200lops kernels.
"""
from __future__ import annotations

from typing import Callable

import numpy as np

from .comm import COMM_WORLD, resolve
from .DistributedArray import DistributedArray

__all__ = ["MPILinearOperator", "asmpilinearoperator"]


def _isintlike(p) -> bool:
    try:
        return int(p) == p and np.ndim(p) == 0
    except (TypeError, ValueError):
        return False


def _get_dtype(operators, dtypes=None):
    """result dtype of a set of operators (scipy's _get_dtype, used at
    LinearOperator.py:279, 307)"""
    dts = list(dtypes or [])
    for op in operators:
        if op is not None and getattr(op, "dtype", None) is not None:
            dts.append(np.dtype(op.dtype) if not isinstance(op.dtype, np.dtype) else op.dtype)
    dts = [np.dtype(d) if not isinstance(d, type) or d not in (int, float, complex)
           else np.dtype(d) for d in dts]
    return np.result_type(*dts) if dts else None


class MPILinearOperator:
    """LinearOperator.py:14-233.  ``Op`` may be any rank-local operator exposing
    ``shape``, ``dtype``, ``_matvec`` / ``_rmatvec`` on device tensors (the role
    ``pylops.LinearOperator`` plays in the reference)."""

    def __init__(self, Op=None, shape=None, dtype=None, base_comm=COMM_WORLD):
        self.Op = None
        if Op is not None:
            self.Op = Op
            dtype = self.Op.dtype if dtype is None else dtype
            shape = self.Op.shape if shape is None else shape
        if shape is not None:
            self.shape = tuple(int(s) for s in shape)
        if dtype is not None:
            self.dtype = dtype
        self.base_comm = resolve(base_comm)
        self.size = self.base_comm.Get_size()
        self.rank = self.base_comm.Get_rank()

    def matvec(self, x: DistributedArray) -> DistributedArray:
        M, N = self.shape
        if x.global_shape != (N,):
            raise ValueError("dimension mismatch")
        return self._matvec(x)

    def _matvec(self, x: DistributedArray) -> DistributedArray:
        if self.Op is not None:
            y = DistributedArray(global_shape=self.shape[0], base_comm=self.base_comm,
                                 partition=x.partition, axis=x.axis, dtype=self.dtype)
            y[:] = self.Op._matvec(x.local_array)
            return y

    def rmatvec(self, x: DistributedArray) -> DistributedArray:
        M, N = self.shape
        if x.global_shape != (M,):
            raise ValueError("dimension mismatch")
        return self._rmatvec(x)

    def _rmatvec(self, x: DistributedArray) -> DistributedArray:
        if self.Op is not None:
            y = DistributedArray(global_shape=self.shape[1], base_comm=self.base_comm,
                                 partition=x.partition, axis=x.axis, dtype=self.dtype)
            y[:] = self.Op._rmatvec(x.local_array)
            return y

    def dot(self, x):
        if isinstance(x, MPILinearOperator):
            return _ProductLinearOperator(self, x)
        elif np.isscalar(x):
            return _ScaledLinearOperator(self, x)
        else:
            if x is None or x.ndim == 1:
                return self.matvec(x)
            raise ValueError('expected 1-d DistributedArray, got %r' % (x.global_shape,))

    def adjoint(self):
        return self._adjoint()

    H = property(adjoint)

    def transpose(self):
        return self._transpose()

    T = property(transpose)

    def __mul__(self, x):
        return self.dot(x)

    def __rmul__(self, x):
        if np.isscalar(x):
            return _ScaledLinearOperator(self, x)
        return NotImplemented

    def __matmul__(self, x):
        if np.isscalar(x):
            raise ValueError("Scalar not allowed, use * instead")
        return self.__mul__(x)

    def __rmatmul__(self, x):
        if np.isscalar(x):
            raise ValueError("Scalar not allowed, use * instead")
        return self.__rmul__(x)

    def __pow__(self, p):
        return _PowerLinearOperator(self, p)

    def __add__(self, x):
        return _SumLinearOperator(self, x)

    def __neg__(self):
        return _ScaledLinearOperator(self, -1)

    def __sub__(self, x):
        return self.__add__(-x)

    def _adjoint(self):
        return _AdjointLinearOperator(self)

    def _transpose(self):
        return _TransposedLinearOperator(self)

    def conj(self):
        return _ConjLinearOperator(self)

    def __repr__(self):
        M, N = self.shape
        dt = "unspecified dtype" if getattr(self, "dtype", None) is None else f"dtype={self.dtype}"
        return f"<{M}x{N} {self.__class__.__name__} with {dt}>"


class _AdjointLinearOperator(MPILinearOperator):
    def __init__(self, A: MPILinearOperator):
        self.A = A
        self.args = (A,)
        super().__init__(shape=(A.shape[1], A.shape[0]), dtype=A.dtype, base_comm=A.base_comm)

    def _matvec(self, x):
        return self.A.rmatvec(x)

    def _rmatvec(self, x):
        return self.A.matvec(x)


class _TransposedLinearOperator(MPILinearOperator):
    def __init__(self, A: MPILinearOperator):
        self.A = A
        self.args = (A,)
        super().__init__(shape=(A.shape[1], A.shape[0]), dtype=A.dtype, base_comm=A.base_comm)

    def _matvec(self, x):
        return self.A.rmatvec(x.conj()).conj()

    def _rmatvec(self, x):
        return self.A.matvec(x.conj()).conj()


class _ProductLinearOperator(MPILinearOperator):
    def __init__(self, A: MPILinearOperator, B: MPILinearOperator):
        if not isinstance(A, MPILinearOperator) or not isinstance(B, MPILinearOperator):
            raise ValueError('both operands have to be a LinearOperator')
        if A.shape[1] != B.shape[0]:
            raise ValueError('cannot multiply %r and %r: shape mismatch' % (A, B))
        self.args = (A, B)
        super().__init__(shape=(A.shape[0], B.shape[1]), dtype=_get_dtype([A, B]),
                         base_comm=A.base_comm)

    def _matvec(self, x):
        return self.args[0].matvec(self.args[1].matvec(x))

    def _rmatvec(self, x):
        return self.args[1].rmatvec(self.args[0].rmatvec(x))

    def _adjoint(self):
        A, B = self.args
        return B.H * A.H


class _ScaledLinearOperator(MPILinearOperator):
    def __init__(self, A: MPILinearOperator, alpha):
        if not isinstance(A, MPILinearOperator):
            raise ValueError('MPILinearOperator expected as A')
        if not np.isscalar(alpha):
            raise ValueError('scalar expected as alpha')
        self.args = (A, alpha)
        super().__init__(shape=A.shape, dtype=_get_dtype([A], [type(alpha)]), base_comm=A.base_comm)

    def _matvec(self, x):
        y = self.args[0].matvec(x)
        if y is not None:
            y.scale_(self.args[1])
        return y

    def _rmatvec(self, x):
        y = self.args[0].rmatvec(x)
        if y is not None:
            y.scale_(np.conj(self.args[1]))
        return y

    def _adjoint(self):
        A, alpha = self.args
        return A.H * np.conj(alpha)


class _SumLinearOperator(MPILinearOperator):
    def __init__(self, A: MPILinearOperator, B: MPILinearOperator):
        if not isinstance(A, MPILinearOperator) or not isinstance(B, MPILinearOperator):
            raise ValueError('both operands have to be a MPILinearOperator')
        if A.shape != B.shape:
            raise ValueError("cannot add %r and %r: shape mismatch" % (A, B))
        self.args = (A, B)
        super().__init__(shape=A.shape, dtype=A.dtype, base_comm=A.base_comm)

    def _matvec(self, x):
        return self.args[0].matvec(x) + self.args[1].matvec(x)

    def _rmatvec(self, x):
        return self.args[0].rmatvec(x) + self.args[1].rmatvec(x)

    def _adjoint(self):
        A, B = self.args
        return A.H + B.H


class _PowerLinearOperator(MPILinearOperator):
    def __init__(self, A: MPILinearOperator, p: int) -> None:
        if not isinstance(A, MPILinearOperator):
            raise ValueError("LinearOperator expected as A")
        if A.shape[0] != A.shape[1]:
            raise ValueError("square LinearOperator expected, got %r" % A)
        if not _isintlike(p) or p < 0:
            raise ValueError("non-negative integer expected as p")
        super().__init__(shape=A.shape, dtype=A.dtype, base_comm=A.base_comm)
        self.args = (A, p)

    def _power(self, fun: Callable, x: DistributedArray) -> DistributedArray:
        res = x.copy()
        for _ in range(self.args[1]):
            res[:] = fun(res).local_array
        return res

    def _matvec(self, x):
        return self._power(self.args[0].matvec, x)

    def _rmatvec(self, x):
        return self._power(self.args[0].rmatvec, x)


class _ConjLinearOperator(MPILinearOperator):
    def __init__(self, A: MPILinearOperator):
        if not isinstance(A, MPILinearOperator):
            raise TypeError('A must be a MPILinearOperator')
        self.A = A
        super().__init__(shape=A.shape, dtype=A.dtype, base_comm=A.base_comm)

    def _matvec(self, x):
        y = self.A.matvec(x.conj())
        return y.conj() if y is not None else y

    def _rmatvec(self, x):
        y = self.A.rmatvec(x.conj())
        return y.conj() if y is not None else y

    def _adjoint(self):
        return _ConjLinearOperator(self.A.H)


def asmpilinearoperator(Op):
    """LinearOperator.py:419-438"""
    if isinstance(Op, MPILinearOperator):
        return Op
    return MPILinearOperator(Op=Op, base_comm=COMM_WORLD)

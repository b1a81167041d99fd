"""``MPIStackedLinearOperator`` and its algebra (pylops_mpi/StackedLinearOperator.py:13-408): operators acting on /
producing :class:`StackedDistributedArray` ("next" row f3) -- composition glue over the B200 operators."""
from __future__ import annotations

import numpy as np

from .comm import COMM_WORLD, resolve
from .DistributedArray import DistributedArray
from .LinearOperator import MPILinearOperator, _get_dtype
from .StackedArray import StackedDistributedArray

__all__ = ["MPIStackedLinearOperator"]


def _stacked_len(x) -> int:
    if isinstance(x, StackedDistributedArray):
        return int(sum(int(np.prod(d.global_shape)) for d in x.distarrays))
    return int(np.prod(x.global_shape))


class MPIStackedLinearOperator:
    """Operators acting on / producing :class:`StackedDistributedArray` (StackedLinearOperator.py:13-228): same
    checks, algebra (``.H .T conj * @ + - ** neg``) and errors as the reference."""

    def __init__(self, shape=None, dtype=None, base_comm=COMM_WORLD):
        self.shape = None if shape is None else (int(shape[0]), int(shape[1]))
        self.dtype = None if dtype is None else np.dtype(dtype)
        self.base_comm = resolve(base_comm)
        self.size, self.rank = self.base_comm.Get_size(), self.base_comm.Get_rank()

    def matvec(self, x):
        if _stacked_len(x) != self.shape[1] or (isinstance(x, DistributedArray) and x.ndim != 1):
            raise ValueError("dimension mismatch")
        return self._matvec(x)

    def rmatvec(self, x):
        if _stacked_len(x) != self.shape[0] or (isinstance(x, DistributedArray) and x.ndim != 1):
            raise ValueError("dimension mismatch")
        return self._rmatvec(x)

    def _matvec(self, x):
        raise NotImplementedError

    def _rmatvec(self, x):
        raise NotImplementedError

    def dot(self, x):
        if isinstance(x, MPIStackedLinearOperator):
            return _ProductStackedLinearOperator(self, x)
        if np.isscalar(x):
            return _ScaledStackedLinearOperator(self, x)
        if isinstance(x, DistributedArray) and x.ndim == 1:
            return self.matvec(x)
        if isinstance(x, StackedDistributedArray) and all(d.ndim == 1 for d in x.distarrays):
            return self.matvec(x)
        raise ValueError("expected 1-d DistributedArray or StackedDistributedArray")

    def adjoint(self):
        return self._adjoint()

    H = property(adjoint)

    def transpose(self):
        return self._transpose()

    T = property(transpose)

    def __mul__(self, x):
        return self.dot(x)

    def __rmul__(self, x):
        return _ScaledStackedLinearOperator(self, x) if np.isscalar(x) else NotImplemented

    def __matmul__(self, x):
        if np.isscalar(x):
            raise ValueError("Scalar not allowed, use * instead")
        return self.__mul__(x)

    def __rmatmul__(self, x):
        if np.isscalar(x):
            raise ValueError("Scalar not allowed, use * instead")
        return self.__rmul__(x)

    def __pow__(self, p):
        return _PowerStackedLinearOperator(self, p)

    def __add__(self, x):
        return _SumStackedLinearOperator(self, x)

    def __neg__(self):
        return _ScaledStackedLinearOperator(self, -1)

    def __sub__(self, x):
        return self.__add__(-x)

    def _adjoint(self):
        return _LazyStacked((self.shape[1], self.shape[0]), self.dtype, self.base_comm, self._rmatvec, self._matvec,
                            lambda: self)

    def _transpose(self):
        return _LazyStacked((self.shape[1], self.shape[0]), self.dtype, self.base_comm,
                            lambda x: self._rmatvec(x.conj()).conj(), lambda x: self._matvec(x.conj()).conj(), None)

    def conj(self):
        return _LazyStacked(self.shape, self.dtype, self.base_comm, lambda x: self.matvec(x.conj()).conj(),
                            lambda x: self.rmatvec(x.conj()).conj(), lambda: self.H.conj())

    def __repr__(self):
        M, N = self.shape
        dt = "unspecified dtype" if self.dtype is None else f"dtype={self.dtype}"
        return f"<{M}x{N} {self.__class__.__name__} with {dt}>"


class _LazyStacked(MPIStackedLinearOperator):
    """adjoint / transpose / conj views (StackedLinearOperator.py:230-266, 383-408) as closures"""

    def __init__(self, shape, dtype, base_comm, mv, rmv, adj):
        super().__init__(shape=shape, dtype=dtype, base_comm=base_comm)
        self._mv, self._rmv, self._adj = mv, rmv, adj

    def _matvec(self, x):
        return self._mv(x)

    def _rmatvec(self, x):
        return self._rmv(x)

    def _adjoint(self):
        return self._adj() if self._adj is not None else super()._adjoint()


class _ProductStackedLinearOperator(MPIStackedLinearOperator):
    """StackedLinearOperator.py:268-293"""

    def __init__(self, A, B):
        from .basicoperators.BlockDiag import MPIStackedBlockDiag      # defined on top of this module: import late
        from .basicoperators.VStack import MPIStackedVStack
        if not isinstance(A, MPIStackedLinearOperator) or not isinstance(B, MPIStackedLinearOperator):
            raise ValueError("both operands have to be a MPIStackedLinearOperator")
        if isinstance(A, MPIStackedVStack) and isinstance(B, MPIStackedVStack):
            raise ValueError("both operands cannot be MPIStackedVStack")
        if isinstance(A, MPIStackedBlockDiag) and isinstance(B, MPIStackedBlockDiag) and len(A.ops) != len(B.ops):
            raise ValueError(f"both MPIStackedBlockDiag cannot have different number of ops, {A.ops} != {B.ops}")
        if A.shape[1] != B.shape[0]:
            raise ValueError("cannot multiply %r and %r: shape mismatch" % (A, B))
        self.args = (A, B)
        super().__init__(shape=(A.shape[0], B.shape[1]), dtype=_get_dtype([A, B]), base_comm=A.base_comm)

    def _matvec(self, x):
        return self.args[0].matvec(self.args[1].matvec(x))

    def _rmatvec(self, x):
        return self.args[1].rmatvec(self.args[0].rmatvec(x))

    def _adjoint(self):
        A, B = self.args
        return B.H * A.H


class _ScaledStackedLinearOperator(MPIStackedLinearOperator):
    """StackedLinearOperator.py:296-323"""

    def __init__(self, A, alpha):
        if not isinstance(A, MPIStackedLinearOperator):
            raise ValueError("MPIStackedLinearOperator expected as A")
        if not np.isscalar(alpha):
            raise ValueError("scalar expected as alpha")
        self.args = (A, alpha)
        super().__init__(shape=A.shape, dtype=_get_dtype([A], [type(alpha)]), base_comm=A.base_comm)

    def _matvec(self, x):
        return self.args[0].matvec(x) * self.args[1]

    def _rmatvec(self, x):
        return self.args[0].rmatvec(x) * np.conj(self.args[1])

    def _adjoint(self):
        A, alpha = self.args
        return A.H * np.conj(alpha)


class _SumStackedLinearOperator(MPIStackedLinearOperator):
    """StackedLinearOperator.py:326-352"""

    def __init__(self, A, B):
        if not isinstance(A, MPIStackedLinearOperator) or not isinstance(B, MPIStackedLinearOperator):
            raise ValueError("both operands have to be a MPIStackedLinearOperator")
        if A.shape != B.shape:
            raise ValueError("cannot add %r and %r: shape mismatch" % (A, B))
        self.args = (A, B)
        super().__init__(shape=A.shape, dtype=_get_dtype([A, B]), base_comm=A.base_comm)

    def _matvec(self, x):
        return self.args[0].matvec(x) + self.args[1].matvec(x)

    def _rmatvec(self, x):
        return self.args[0].rmatvec(x) + self.args[1].rmatvec(x)

    def _adjoint(self):
        A, B = self.args
        return A.H + B.H


class _PowerStackedLinearOperator(MPIStackedLinearOperator):
    """StackedLinearOperator.py:355-380"""

    def __init__(self, A, p):
        if not isinstance(A, MPIStackedLinearOperator):
            raise ValueError("MPIStackedLinearOperator expected as A")
        if A.shape[0] != A.shape[1]:
            raise ValueError("square MPIStackedLinearOperator expected, got %r" % A)
        if not isinstance(p, (int, np.integer)) or p < 0:
            raise ValueError("non-negative integer expected as p")
        super().__init__(shape=A.shape, dtype=A.dtype, base_comm=A.base_comm)
        self.args = (A, int(p))

    def _power(self, fun, x):
        res = x.copy()
        for _ in range(self.args[1]):
            res = fun(res)
        return res

    def _matvec(self, x):
        return self._power(self.args[0].matvec, x)

    def _rmatvec(self, x):
        return self._power(self.args[0].rmatvec, x)

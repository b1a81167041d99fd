"""Minimal stand-in for the parts of third-party ``pylops`` that the reference's
hot-path modules import (TEST INFRASTRUCTURE for tests/golden/make_golden.py).
``MatrixMult`` is the dense block ``A @ x`` / ``A^H @ x`` -- the only third-party
arithmetic on this path."""
import numpy as np


class LinearOperator:
    def __init__(self, dtype=None, shape=None):
        self.dtype = dtype
        self.shape = shape

    def matvec(self, x):
        return self._matvec(x)

    def rmatvec(self, x):
        return self._rmatvec(x)


class MatrixMult(LinearOperator):
    def __init__(self, A, dtype="float64"):
        self.A = A
        super().__init__(dtype=np.dtype(dtype), shape=A.shape)

    def _matvec(self, x):
        return self.A @ x

    def _rmatvec(self, x):
        return self.A.conj().T @ x


class Identity(LinearOperator):
    """pylops.Identity(N, M): keep the first N of M samples (adjoint: zero-pad) -- restated third-party operator"""

    def __init__(self, N, M=None, inplace=True, dtype="float64"):
        M = N if M is None else M
        self.inplace = inplace
        super().__init__(dtype=np.dtype(dtype), shape=(int(N), int(M)))

    def _matvec(self, x):
        N, M = self.shape
        if N <= M:
            return x[:N] if self.inplace else x[:N].copy()
        y = np.zeros(N, dtype=self.dtype)
        y[:M] = x
        return y

    def _rmatvec(self, x):
        N, M = self.shape
        if M <= N:
            return x[:M] if self.inplace else x[:M].copy()
        y = np.zeros(M, dtype=self.dtype)
        y[:N] = x
        return y


from . import utils, optimization  # noqa: E402,F401
from ._derivatives import FirstDerivative, SecondDerivative  # noqa: E402,F401
from . import basicoperators, signalprocessing  # noqa: E402,F401

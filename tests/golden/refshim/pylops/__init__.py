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


from . import utils, optimization  # noqa: E402,F401
from ._derivatives import FirstDerivative, SecondDerivative  # noqa: E402,F401
from . import basicoperators  # noqa: E402,F401

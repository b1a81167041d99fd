"""Import-path parity with ``pylops_mpi/StackedLinearOperator.py``; the implementation lives in StackedArray.py."""
from .StackedArray import MPIStackedLinearOperator  # noqa: F401

__all__ = ["MPIStackedLinearOperator"]

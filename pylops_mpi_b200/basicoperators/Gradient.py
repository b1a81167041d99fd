"""Import-path parity with ``pylops_mpi/basicoperators/Gradient.py``; the implementation lives in StackedArray.py."""
from ..StackedArray import MPIGradient  # noqa: F401

__all__ = ["MPIGradient"]

from .Fredholm1 import MPIFredholm1  # noqa: F401

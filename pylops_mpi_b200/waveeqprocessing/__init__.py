from .MDC import MPIMDC  # noqa: F401

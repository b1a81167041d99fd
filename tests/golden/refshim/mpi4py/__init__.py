"""In-process stand-in for mpi4py (TEST INFRASTRUCTURE, used only by
tests/golden/make_golden.py to import and run the REAL reference from
/root/reference without an MPI installation).  Each simulated rank is a Python
thread; communicators share slots guarded by barriers; point-to-point messages
go through queues.  Only the calls the reference's hot path makes are provided.
"""
from . import MPI  # noqa: F401

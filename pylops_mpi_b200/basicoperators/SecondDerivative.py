"""``MPISecondDerivative`` (pylops_mpi/basicoperators/SecondDerivative.py:13-257): second derivative
along axis 0 of a row-block distributed model -- forward / backward / centered 3-point stencils,
optional edges, exact-transpose adjoint.  Same single-kernel + single-halo-exchange apply as
:class:`MPIFirstDerivative` (the reference does three ghost-cell exchanges per adjoint, :221-246)."""
from __future__ import annotations

import ctypes as C

import numpy as np

from .. import _lib
from ..comm import COMM_WORLD
from .FirstDerivative import MPIFirstDerivative, _KINDS


class MPISecondDerivative(MPIFirstDerivative):
    _deriv = 2

    def __init__(self, dims, sampling: float = 1.0, kind: str = "centered", edge: bool = False,
                 base_comm=COMM_WORLD, dtype=np.float64):
        if kind not in _KINDS:
            raise NotImplementedError("'kind' must be 'forward', 'centered' or 'backward'")
        super().__init__(dims, sampling=sampling, kind=kind, edge=edge, order=3, base_comm=base_comm, dtype=dtype)

    def _halo_need(self, adjoint: bool):
        need_lo, need_hi = C.c_int(), C.c_int()
        _lib.check(_lib.lib.b2_second_derivative_halo(self._kind_code, int(self.edge), int(adjoint),
                                                      C.byref(need_lo), C.byref(need_hi)), "b2_second_derivative_halo")
        return need_lo.value, need_hi.value

    def _kernel(self, ctx, xp, yp, lop, lo_n, hip, hi_n, nrows, ncols, row0, adjoint, code):
        _lib.check(_lib.lib.b2_second_derivative(ctx, xp, yp, lop, lo_n, hip, hi_n, nrows, ncols, row0, self.dims[0],
                                                 self._kind_code, int(self.edge), float(self.sampling), adjoint, code,
                                                 _lib.stream()), "b2_second_derivative")

"""Adjoint (dot) test, pylops_mpi/utils/dottest.py:11-107."""
from __future__ import annotations

from typing import Optional

import numpy as np


def dottest(Op, u, v, nr: Optional[int] = None, nc: Optional[int] = None, rtol: float = 1e-6,
            atol: float = 1e-21, raiseerror: bool = True, verb: bool = False) -> bool:
    if nr is None:
        nr = Op.shape[0]
    if nc is None:
        nc = Op.shape[1]
    if (nr, nc) != Op.shape:
        raise AssertionError("Provided nr and nc do not match operator shape")
    y = Op.matvec(u)   # Op * u
    x = Op.rmatvec(v)  # Op'* v
    yy = np.vdot(y.asarray().cpu().numpy(), v.asarray().cpu().numpy())  # (Op  * u)' * v
    xx = np.vdot(u.asarray().cpu().numpy(), x.asarray().cpu().numpy())  # u' * (Op' * v)
    passed = bool(np.isclose(xx, yy, rtol, atol))
    if (not passed and raiseerror) or verb:
        passed_status = "passed" if passed else "failed"
        msg = f"Dot test {passed_status}, v^H(Opu)={yy} - u^H(Op^Hv)={xx}"
        if not passed and raiseerror:
            raise AssertionError(msg)
        print(msg)
    return passed

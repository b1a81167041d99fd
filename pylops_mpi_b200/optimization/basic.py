"""Functional wrappers ``cg`` / ``cgls`` (pylops_mpi/optimization/basic.py:13-148)."""
from __future__ import annotations

from typing import Callable, Optional, Tuple

from ..DistributedArray import DistributedArray
from .cls_basic import CG, CGLS


def cg(Op, y, x0: Optional[DistributedArray] = None, niter: int = 10, tol: float = 1e-4,
       show: bool = False, itershow: Tuple[int, int, int] = (10, 10, 10),
       callback: Optional[Callable] = None):
    cgsolve = CG(Op)
    if callback is not None:
        cgsolve.callback = callback
    x, iiter, cost = cgsolve.solve(y=y, x0=x0, tol=tol, niter=niter, show=show, itershow=itershow)
    return x, iiter, cost


def cgls(Op, y, x0: Optional[DistributedArray] = None, niter: int = 10, damp: float = 0.0,
         tol: float = 1e-4, show: bool = False, itershow: Tuple[int, int, int] = (10, 10, 10),
         callback: Optional[Callable] = None):
    cgsolve = CGLS(Op)
    if callback is not None:
        cgsolve.callback = callback
    x, istop, iiter, r1norm, r2norm, cost = cgsolve.solve(y=y, x0=x0, niter=niter, damp=damp,
                                                          tol=tol, show=show, itershow=itershow)
    return x, istop, iiter, r1norm, r2norm, cost

"""Functional wrappers ``ista`` / ``fista`` (pylops_mpi/optimization/sparsity.py:11-257).
``rtol`` / ``rtol1`` are accepted for signature compatibility: in the reference they only attach third-party
pylops callbacks (CostToInitialCallback / CostToDataCallback) that its own run loops never consult."""
from typing import Any, Callable, Dict, Optional, Tuple

from .cls_sparsity import FISTA, ISTA


def _solve(cls, Op, y, x0, niter, SOp, eps, alpha, eigsdict, tol, threshkind, decay, monitorres, show, itershow,
           callback):
    solver = cls(Op)
    if callback is not None:
        solver.callback = callback
    return solver.solve(y=y, x0=x0, niter=niter, SOp=SOp, eps=eps, alpha=alpha, eigsdict=eigsdict, tol=tol,
                        threshkind=threshkind, decay=decay, monitorres=monitorres, show=show, itershow=itershow)


def ista(Op, y, x0, niter: int = 10, SOp=None, eps: float = 0.1, alpha: Optional[float] = None,
         eigsdict: Optional[Dict[str, Any]] = None, tol: float = 1e-10, rtol: float = 0.0, rtol1: float = 0.0,
         threshkind: str = "soft", decay=None, monitorres: bool = False, show: bool = False,
         itershow: Tuple[int, int, int] = (10, 10, 10), callback: Optional[Callable] = None):
    return _solve(ISTA, Op, y, x0, niter, SOp, eps, alpha, eigsdict, tol, threshkind, decay, monitorres, show,
                  itershow, callback)


def fista(Op, y, x0, niter: int = 10, SOp=None, eps: float = 0.1, alpha: Optional[float] = None,
          eigsdict: Optional[Dict[str, Any]] = None, tol: float = 1e-10, rtol: float = 0.0, rtol1: float = 0.0,
          threshkind: str = "soft", decay=None, monitorres: bool = False, show: bool = False,
          itershow: Tuple[int, int, int] = (10, 10, 10), callback: Optional[Callable] = None):
    return _solve(FISTA, Op, y, x0, niter, SOp, eps, alpha, eigsdict, tol, threshkind, decay, monitorres, show,
                  itershow, callback)

"""ISTA / FISTA with the reference's recurrences, stopping rule and outputs
(pylops_mpi/optimization/cls_sparsity.py:49-715) -- "next" row of SURVEY section 8f.

Two execution modes, same numbers:
  * generic (``SOp`` given, stacked or BROADCAST models): the reference's sequence of DistributedArray
    operations, the threshold itself being one in-place CUDA pass (``b2_sparse_update`` without gradient);
  * fused (default: SCATTER DistributedArray model, no ``SOp``): gradient step, threshold, FISTA momentum,
    update norm and l1 cost in ONE pass over the model (``b2_sparse_update``), and the three scalars of the
    iteration (|x - xold|^2, |x|_1, |res|^2) in one all-reduce + one host read.
"""
from __future__ import annotations

import time
from math import sqrt
from typing import Any, Dict, Optional, Tuple

import numpy as np
import torch

from .. import _lib
from ..Distributed import allreduce_
from ..DistributedArray import DistributedArray, Partition
from ..StackedArray import StackedDistributedArray
from .cls_basic import Solver, _dots_device
from .eigs import power_iteration

_KINDS = {"soft": _lib.THRESH_SOFT, "hard": _lib.THRESH_HARD, "half": _lib.THRESH_HALF}


def _sparse_update(base: torch.Tensor, g: Optional[torch.Tensor], alpha: float, xold: Optional[torch.Tensor],
                   thresh: float, kind: int, xnew: torch.Tensor, znew: Optional[torch.Tensor], c: float,
                   sums: Optional[torch.Tensor]):
    n = base.numel()
    ptr = lambda t: t.data_ptr() if (t is not None and n) else None   # noqa: E731
    _lib.check(_lib.lib.b2_sparse_update(_lib.ctx(), ptr(base), ptr(g), float(alpha), ptr(xold), float(thresh), kind,
                                         ptr(xnew), ptr(znew), float(c),
                                         sums.data_ptr() if sums is not None else None, n, _lib.code(base.dtype),
                                         _lib.stream()), "b2_sparse_update")


def _apply_thresh(x, kind: int, thresh: float):
    """in-place thresholding of a (stacked) distributed array (cls_sparsity.py:21-46)"""
    for d in (x.distarrays if isinstance(x, StackedDistributedArray) else [x]):
        a = d._cont()
        _sparse_update(a, None, 0.0, None, thresh, kind, a, None, 0.0, None)
        if a is not d.local_array:
            d.local_array.copy_(a)
    return x


class ISTA(Solver):
    """cls_sparsity.py:49-483"""

    def setup(self, y, x0, niter: Optional[int] = None, SOp=None, eps: float = 0.1, alpha: Optional[float] = None,
              eigsdict: Optional[Dict[str, Any]] = None, tol: float = 1e-10, threshkind: str = "soft",
              decay=None, monitorres: bool = False, show: bool = False):
        self.y, self.SOp, self.niter, self.eps = y, SOp, niter, eps
        self.eigsdict = {} if eigsdict is None else eigsdict
        self.tol, self.threshkind, self.decay, self.monitorres = tol, threshkind, decay, monitorres
        if threshkind not in _KINDS:
            raise ValueError(f"threshkind must be hard, soft, half, got {threshkind}")
        self.kind = _KINDS[threshkind]
        if decay is None:
            self.decay = np.ones(niter, dtype=np.empty(0, dtype=self.Op.dtype).real.dtype)
        if alpha is not None:
            self.alpha = alpha
        elif not hasattr(self, "alpha"):
            Op1 = self.Op.H @ self.Op                                            # :230-243
            maxeig = np.abs(power_iteration(Op1, b_k=x0.empty_like(), dtype=Op1.dtype, **self.eigsdict)[0])
            self.alpha = float(1.0 / maxeig)
        self.thresh = eps * self.alpha * 0.5
        x = x0.copy()
        if monitorres:
            self.normresold = np.inf
        self.t = 1.0
        self.cost = []
        self.iiter = 0
        self._fused = (SOp is None and isinstance(x, DistributedArray) and isinstance(y, DistributedArray)
                       and x.partition is Partition.SCATTER and y.partition is Partition.SCATTER)
        if self._fused:
            self._scal = torch.zeros(4, dtype=torch.float64, device=x.local_array.device)
        if show:
            self._print_solver(f"\neps = {eps:10e}\ttol = {tol:10e}\tniter = {niter}\nalpha = {self.alpha:10e}"
                               f"\tthresh = {self.thresh:10e}")
        return x

    # ---- shared pieces ------------------------------------------------------------------------------
    def _residual(self, x):
        res = self.y - self.Op.matvec(x)
        return res

    def _check_res(self, normres: float, name: str):
        if normres > self.normresold:
            raise ValueError(f"{name} stopped at iteration {self.iiter} due to residual increasing, consider "
                             "modifying eps and/or alpha...")
        self.normresold = normres

    def _threshold_generic(self, x_unthesh):
        t = self.decay[self.iiter] * self.thresh
        if self.SOp is None:
            return _apply_thresh(x_unthesh, self.kind, t)
        return self.SOp.matvec(_apply_thresh(self.SOp.rmatvec(x_unthesh), self.kind, t))

    def _fused_scalars(self, res: Optional[DistributedArray]) -> np.ndarray:
        """all-reduce [|dx|^2, |x|_1, |res|^2] together, one host read"""
        if res is not None:
            _dots_device([res], self._scal, 2)
        allreduce_(self.y._sub_comm, self._scal, "sum")
        return self._scal.cpu().numpy()

    def step(self, x, show: bool = False):
        if not self._fused:
            xold = x.copy()
            res = self._residual(x)
            if self.monitorres:
                self._check_res(float(res.norm()[0]), "ISTA")
            x_unthesh = x + self.Op.rmatvec(res) * self.alpha
            x = self._threshold_generic(x_unthesh)
            xupdate = float((x - xold).norm()[0])
            costdata = 0.5 * float(res.norm()[0]) ** 2
            costreg = self.eps * float(x.norm(ord=1)[0])
        else:
            res = self._residual(x)
            g = self.Op.rmatvec(res)
            x._check_partition_shape(g)
            xa = x._cont()
            _sparse_update(xa, g._cont(), self.alpha, xa, self.decay[self.iiter] * self.thresh, self.kind, xa, None,
                           0.0, self._scal)
            if xa is not x.local_array:
                x.local_array.copy_(xa)
            s = self._fused_scalars(res)
            if self.monitorres:
                self._check_res(float(np.sqrt(s[2])), "ISTA")
            xupdate, costdata, costreg = float(np.sqrt(s[0])), 0.5 * float(s[2]), self.eps * float(s[1])
        self.cost.append(float(costdata + costreg))
        self.iiter += 1
        if show:
            print(f"{self.iiter:6g}  {costdata:11.4e}  {costdata + costreg:11.4e}  {xupdate:11.4e}")
        return x, xupdate

    def run(self, x, niter: Optional[int] = None, show: bool = False, itershow: Tuple[int, int, int] = (10, 10, 10)):
        xupdate = np.inf
        niter = self.niter if niter is None else niter
        if niter is None:
            raise ValueError("niter must not be None")
        while self.iiter < niter and xupdate > self.tol:
            showstep = bool(show and (self.iiter < itershow[0] or niter - self.iiter < itershow[1]
                                      or self.iiter % itershow[2] == 0))
            x, xupdate = self.step(x, showstep)
            self.callback(x)
        return x

    def finalize(self, show: bool = False) -> None:
        self.tend = time.time()
        self.telapsed = self.tend - self.tstart
        self.cost = np.array(self.cost)
        if show:
            self._print_finalize()

    def solve(self, y, x0, niter: Optional[int] = None, SOp=None, eps: float = 0.1, alpha: Optional[float] = None,
              eigsdict: Optional[Dict[str, Any]] = None, tol: float = 1e-10, threshkind: str = "soft", decay=None,
              monitorres: bool = False, show: bool = False, itershow: Tuple[int, int, int] = (10, 10, 10)):
        x = self.setup(y=y, x0=x0, niter=niter, SOp=SOp, eps=eps, alpha=alpha, eigsdict=eigsdict, tol=tol,
                       threshkind=threshkind, decay=decay, monitorres=monitorres, show=show)
        x = self.run(x, niter, show=show, itershow=itershow)
        self.finalize(show)
        return x, self.iiter, self.cost


class FISTA(ISTA):
    """cls_sparsity.py:486-715"""

    def step(self, x, z, show: bool = False):
        told = self.t
        t = (1.0 + sqrt(1.0 + 4.0 * told ** 2)) / 2.0
        c = (told - 1.0) / t
        if not self._fused:
            xold = x.copy()
            res = self._residual(z)
            if self.monitorres:
                self._check_res(float(res.norm()[0]), "FISTA")
            x_unthesh = z + self.Op.rmatvec(res) * self.alpha
            x = self._threshold_generic(x_unthesh)
            dx = x - xold
            z = x + dx * c
            xupdate = float(dx.norm()[0])
            costdata = 0.5 * float((self.y - self.Op.matvec(x)).norm()[0]) ** 2
            costreg = self.eps * float(x.norm(ord=1)[0])
        else:
            res = self._residual(z)
            if self.monitorres:
                self._check_res(float(res.norm()[0]), "FISTA")
            g = self.Op.rmatvec(res)
            z._check_partition_shape(g)
            xa, za = x._cont(), z._cont()
            _sparse_update(za, g._cont(), self.alpha, xa, self.decay[self.iiter] * self.thresh, self.kind, xa, za, c,
                           self._scal)
            if xa is not x.local_array:
                x.local_array.copy_(xa)
            if za is not z.local_array:
                z.local_array.copy_(za)
            s = self._fused_scalars(self._residual(x))                           # cost on the NEW x (:652)
            xupdate, costdata, costreg = float(np.sqrt(s[0])), 0.5 * float(s[2]), self.eps * float(s[1])
        self.t = t
        self.cost.append(float(costdata + costreg))
        self.iiter += 1
        if show:
            print(f"{self.iiter:6g}  {costdata:11.4e}  {costdata + costreg:11.4e}  {xupdate:11.4e}")
        return x, z, xupdate

    def run(self, x, niter: Optional[int] = None, show: bool = False, itershow: Tuple[int, int, int] = (10, 10, 10)):
        z = x.copy()
        xupdate = np.inf
        niter = self.niter if niter is None else niter
        if niter is None:
            raise ValueError("niter must not be None")
        while self.iiter < niter and xupdate > self.tol:
            showstep = bool(show and (self.iiter < itershow[0] or niter - self.iiter < itershow[1]
                                      or self.iiter % itershow[2] == 0))
            x, z, xupdate = self.step(x, z, showstep)
            self.callback(x)
        return x

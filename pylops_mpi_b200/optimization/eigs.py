"""Power iteration for the largest eigenvalue (pylops_mpi/optimization/eigs.py:10-98)."""
from __future__ import annotations

import numpy as np
import torch

from ..DistributedArray import DistributedArray
from ..StackedArray import StackedDistributedArray


def _randomize(d: DistributedArray, dtype, cmpx):
    n = d.local_shape[0]
    v = np.random.rand(n).astype(dtype) + cmpx * np.random.rand(n).astype(dtype)  # host RNG, as eigs.py:62-72
    d[:] = torch.as_tensor(v).to(d.local_array.device)


def power_iteration(Op, niter: int = 10, tol: float = 1e-5, dtype="float32", backend: str = "b200", b_k=None):
    dtype = np.dtype(dtype)
    cmpx = 1j if np.issubdtype(dtype, np.complexfloating) else 0
    stacked = isinstance(b_k, StackedDistributedArray)
    if b_k is None:
        b_k = DistributedArray(global_shape=Op.shape[1], base_comm=Op.base_comm, dtype=dtype)
    for d in (b_k.distarrays if stacked else [b_k]):
        _randomize(d, dtype, cmpx)
    inv = 1.0 / float(b_k.norm()[0])
    for d in (b_k.distarrays if stacked else [b_k]):
        d.scale_(inv)
    maxeig_old = 0.0
    iiter = -1
    for iiter in range(niter):
        b1_k = Op.matvec(b_k)
        maxeig = b_k.dot(b1_k, vdot=True).item()
        inv = 1.0 / float(b1_k.norm()[0])
        for d, d1 in zip(b_k.distarrays if stacked else [b_k], b1_k.distarrays if stacked else [b1_k]):
            d[:] = d1.local_array
            d.scale_(inv)
        if np.abs(maxeig - maxeig_old) < tol * maxeig:
            break
        maxeig_old = maxeig
    return maxeig, b_k, iiter + 1

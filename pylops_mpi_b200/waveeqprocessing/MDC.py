"""``MPIMDC`` -- multi-dimensional convolution ``F1^H I1^H Fredholm1 I F``
(pylops_mpi/waveeqprocessing/MDC.py:12-180).  The distributed, compute-heavy stage is
:class:`MPIFredholm1` (fused product + all-gather kernel); the FFT / frequency-truncation stages are
rank-replicated local operators (``local.FFT``, ``local.Identity``) wrapped in ``MPILinearOperator``,
exactly as the reference composes third-party pylops operators.  The reference has no MDC test and
pylops is not available here: the FFT stage's convention is restated (see ``local.FFT``); the pipeline is
checked against fixtures produced by the reference's own chain run over that restatement in NumPy
(tests/golden ``mdc/``), the oracle and a dot-test."""
from __future__ import annotations

import logging

import numpy as np
import torch

from ..comm import COMM_WORLD
from ..LinearOperator import MPILinearOperator
from ..local import FFT, Identity
from ..signalprocessing.Fredholm1 import MPIFredholm1


def MPIMDC(G, nt: int, nv: int, nfreq: int, dt: float = 1.0, dr: float = 1.0, twosided: bool = True,
           fftengine: str = "numpy", saveGt: bool = True, conj: bool = False, usematmul: bool = False,
           prescaled: bool = False, base_comm=COMM_WORLD, data_domain: str = "time"):
    """Same signature as the reference (MDC.py:77-90); ``G`` is this rank's batch of frequency slices
    ``(nfreq_rank, ns, nr)`` (complex); ``fftengine`` is accepted and ignored.

    ``data_domain="frequency"`` (extension): the operator stops after the Fredholm stage,
    ``Fredholm1 I F`` -- the data side is the band-limited SPECTRUM, kept SCATTERed over the ranks by frequency
    slice.  The forward apply then needs NO Allgather at all and the adjoint a single one (of the model-side
    spectrum), instead of one 64 MiB gather in each direction.  Because ``F1^H I1^H`` has orthonormal columns for
    physical kernels (real DC slice), CGLS on ``|| I1 F1 d - Fredholm1 I F m ||`` produces the same iterates as on
    the time-domain residual; :func:`mdc_data_to_frequency` maps the time-domain data once."""
    if data_domain not in ("time", "frequency"):
        raise ValueError("data_domain must be 'time' or 'frequency'")
    if twosided and nt % 2 == 0:
        raise ValueError('nt must be odd number')
    if not isinstance(G, torch.Tensor):
        G = torch.as_tensor(np.asarray(G))
    cdtype = G.dtype
    rdtype = {torch.complex64: torch.float32, torch.complex128: torch.float64}.get(cdtype, cdtype)
    nfmax = nfreq
    Gs = G if prescaled else (dr * dt * np.sqrt(nt)) * G
    Frop = MPIFredholm1(Gs, nv, saveGt=saveGt, usematmul=usematmul, base_comm=base_comm, dtype=_np_of(cdtype),
                        scatter_data=(data_domain == "frequency"))
    Fr0 = Frop            # the MPIFredholm1 itself (slice bookkeeping), also when wrapped by .conj()
    if conj:
        Frop = Frop.conj()
    _, ns, nr = G.shape
    nfft = int(np.ceil((nt + 1) / 2))
    if nfmax > nfft:
        nfmax = nfft
        logging.warning('nfmax set equal to ceil[(nt+1)/2=%d]' % nfmax)
    Fop = MPILinearOperator(FFT(dims=(nt, nr, nv), axis=0, real=True, ifftshift_before=twosided, dtype=rdtype),
                            base_comm=base_comm)
    F1op = MPILinearOperator(FFT(dims=(nt, ns, nv), axis=0, real=True, ifftshift_before=False, dtype=rdtype),
                             base_comm=base_comm)
    Iop = MPILinearOperator(Identity(N=nfmax * nr * nv, M=nfft * nr * nv, dtype=_np_of(cdtype)), base_comm=base_comm)
    I1op = MPILinearOperator(Identity(N=nfmax * ns * nv, M=nfft * ns * nv, dtype=_np_of(cdtype)), base_comm=base_comm)
    if data_domain == "frequency":
        MDCop = Frop * Iop * Fop
        MDCop.data_to_frequency = lambda d: mdc_data_to_frequency(d, Fr0, I1op, F1op)
        MDCop.dtype = np.dtype(_np_of(cdtype))
        return MDCop
    MDCop = F1op.H * I1op.H * Frop * Iop * Fop
    MDCop.dtype = np.dtype(_np_of(rdtype))   # as the reference: labelled real, carried as complex arrays
    return MDCop


def mdc_data_to_frequency(d, Frop, I1op, F1op):
    """``I1 F1 d`` restricted to this rank's frequency slices: the time-domain (BROADCAST) data of an MDD problem
    mapped ONCE to the SCATTERed band-limited spectrum the ``data_domain="frequency"`` operator works on"""
    from ..DistributedArray import DistributedArray, Partition
    spec = (I1op * F1op).matvec(d)                  # BROADCAST, nfmax * ns * nv
    per = Frop.nx * Frop.nz
    rank = Frop.rank
    out = DistributedArray(global_shape=Frop.shape[0], base_comm=d.base_comm, partition=Partition.SCATTER,
                           local_shapes=[(int(n) * per,) for n in Frop.nsls], dtype=Frop._tdtype)
    src = spec.local_array.reshape(-1)[int(Frop.islstart[rank]) * per: int(Frop.islend[rank]) * per]
    out.local_array.copy_(src)
    return out


def _np_of(t):
    from .. import _lib
    return _lib.numpy_dtype(t)

"""Restatement of third-party ``pylops.signalprocessing.FFT`` (pylops 2.x ``_FFT_numpy``: engine numpy,
norm="ortho", real one-sided transform with the sqrt(2) scaling of the positive frequencies that makes the
adjoint exact) -- TEST INFRASTRUCTURE so that the reference's MPIMDC glue (waveeqprocessing/MDC.py) can be run
by tests/golden/make_golden.py.  ``Fredholm1`` is only a default argument there (MPIMDC passes MPIFredholm1)."""
import numpy as np

from .. import LinearOperator


class FFT(LinearOperator):
    def __init__(self, dims, axis=-1, nfft=None, sampling=1.0, norm="ortho", real=False, ifftshift_before=False,
                 fftshift_after=False, engine="numpy", dtype="complex128"):
        assert norm == "ortho" and engine == "numpy" and not fftshift_after
        self.dims = tuple(int(d) for d in dims)
        self.axis = axis % len(self.dims)
        self.nfft = self.dims[self.axis] if nfft is None else int(nfft)
        self.real, self.ifftshift_before = real, ifftshift_before
        nfo = self.nfft // 2 + 1 if real else self.nfft
        self.dimsd = self.dims[:self.axis] + (nfo,) + self.dims[self.axis + 1:]
        self.rdtype = np.real(np.ones(1, dtype)).dtype
        self.cdtype = (np.ones(1, dtype=self.rdtype) + 1j * np.ones(1, dtype=self.rdtype)).dtype
        self.clinear = False if real else True
        super().__init__(dtype=self.cdtype, shape=(int(np.prod(self.dimsd)), int(np.prod(self.dims))))

    def _matvec(self, x):
        x = np.reshape(x, self.dims)
        if self.ifftshift_before:
            x = np.fft.ifftshift(x, axes=self.axis)
        if not self.clinear:
            x = np.real(x)
        if self.real:
            y = np.fft.rfft(x, n=self.nfft, axis=self.axis, norm="ortho")
            y = np.swapaxes(y, -1, self.axis)
            y[..., 1:1 + (self.nfft - 1) // 2] *= np.sqrt(2)
            y = np.swapaxes(y, self.axis, -1)
        else:
            y = np.fft.fft(x, n=self.nfft, axis=self.axis, norm="ortho")
        return y.astype(self.cdtype).ravel()

    def _rmatvec(self, x):
        x = np.reshape(x, self.dimsd)
        if self.real:
            x = x.copy()
            x = np.swapaxes(x, -1, self.axis)
            x[..., 1:1 + (self.nfft - 1) // 2] /= np.sqrt(2)
            x = np.swapaxes(x, self.axis, -1)
            y = np.fft.irfft(x, n=self.nfft, axis=self.axis, norm="ortho")
        else:
            y = np.fft.ifft(x, n=self.nfft, axis=self.axis, norm="ortho")
        if self.nfft > self.dims[self.axis]:
            y = np.take(y, range(0, self.dims[self.axis]), axis=self.axis)
        if not self.clinear:
            y = np.real(y)
        if self.ifftshift_before:
            y = np.fft.fftshift(y, axes=self.axis)
        return y.astype(self.rdtype if not self.clinear else self.cdtype).ravel()


class Fredholm1:
    def __init__(self, *a, **k):
        raise NotImplementedError("serial pylops.signalprocessing.Fredholm1 is not on this path")

"""Restatement of third-party ``pylops.FirstDerivative`` / ``pylops.SecondDerivative`` (pylops 2.x,
basicoperators/firstderivative.py / secondderivative.py: slice-based stencils along one axis of an
N-d array) -- TEST INFRASTRUCTURE, used only so that the reference's MPIGradient / MPILaplacian glue can be
run by tests/golden/make_golden.py.  Third-party pylops itself is absent from the image."""
import numpy as np

from . import LinearOperator


class _AxisOp(LinearOperator):
    def __init__(self, dims, axis, sampling, kind, edge, dtype):
        self.dims = tuple(int(d) for d in (dims if np.ndim(dims) else (dims,)))
        self.axis = axis % len(self.dims)
        self.sampling, self.kind, self.edge = sampling, kind, edge
        n = int(np.prod(self.dims))
        super().__init__(dtype=np.dtype(dtype), shape=(n, n))

    def _apply(self, x, fn):
        X = np.moveaxis(np.reshape(x, self.dims), self.axis, 0)
        Y = np.zeros(X.shape, dtype=np.result_type(X.dtype, self.dtype))
        fn(X, Y)
        return np.moveaxis(Y, 0, self.axis).ravel()

    def _matvec(self, x):
        return self._apply(x, self._fwd)

    def _rmatvec(self, x):
        return self._apply(x, self._adj)


class FirstDerivative(_AxisOp):
    def __init__(self, dims, axis=-1, sampling=1.0, kind="centered", edge=False, order=3, dtype="float64"):
        super().__init__(dims, axis, sampling, kind, edge, dtype)
        self.order = order

    def _fwd(self, x, y):
        if self.kind == "forward":
            y[:-1] = x[1:] - x[:-1]
        elif self.kind == "backward":
            y[1:] = x[1:] - x[:-1]
        elif self.order == 3:
            y[1:-1] = 0.5 * (x[2:] - x[:-2])
            if self.edge:
                y[0] = x[1] - x[0]
                y[-1] = x[-1] - x[-2]
        else:
            y[2:-2] = x[:-4] / 12.0 - 2 * x[1:-3] / 3.0 + 2 * x[3:-1] / 3.0 - x[4:] / 12.0
            if self.edge:
                y[0] = x[1] - x[0]
                y[1] = 0.5 * (x[2] - x[0])
                y[-2] = 0.5 * (x[-1] - x[-3])
                y[-1] = x[-1] - x[-2]
        y /= self.sampling

    def _adj(self, x, y):
        if self.kind == "forward":
            y[:-1] -= x[:-1]
            y[1:] += x[:-1]
        elif self.kind == "backward":
            y[:-1] -= x[1:]
            y[1:] += x[1:]
        elif self.order == 3:
            y[:-2] -= 0.5 * x[1:-1]
            y[2:] += 0.5 * x[1:-1]
            if self.edge:
                y[0] -= x[0]
                y[1] += x[0]
                y[-2] -= x[-1]
                y[-1] += x[-1]
        else:
            y[:-4] += x[2:-2] / 12.0
            y[1:-3] -= 2.0 * x[2:-2] / 3.0
            y[3:-1] += 2.0 * x[2:-2] / 3.0
            y[4:] -= x[2:-2] / 12.0
            if self.edge:
                y[0] -= x[0] + 0.5 * x[1]
                y[1] += x[0]
                y[2] += 0.5 * x[1]
                y[-3] -= 0.5 * x[-2]
                y[-2] -= x[-1]
                y[-1] += 0.5 * x[-2] + x[-1]
        y /= self.sampling


class SecondDerivative(_AxisOp):
    def __init__(self, dims, axis=-1, sampling=1.0, kind="centered", edge=False, dtype="float64"):
        super().__init__(dims, axis, sampling, kind, edge, dtype)

    def _fwd(self, x, y):
        d = x[2:] - 2 * x[1:-1] + x[:-2]
        if self.kind == "forward":
            y[:-2] = d
        elif self.kind == "backward":
            y[2:] = d
        else:
            y[1:-1] = d
            if self.edge:
                y[0] = x[0] - 2 * x[1] + x[2]
                y[-1] = x[-3] - 2 * x[-2] + x[-1]
        y /= self.sampling ** 2

    def _adj(self, x, y):
        xs = {"forward": x[:-2], "backward": x[2:], "centered": x[1:-1]}[self.kind]
        y[:-2] += xs
        y[1:-1] -= 2 * xs
        y[2:] += xs
        if self.kind == "centered" and self.edge:
            y[0] += x[0]
            y[1] -= 2 * x[0]
            y[2] += x[0]
            y[-3] += x[-1]
            y[-2] -= 2 * x[-1]
            y[-1] += x[-1]
        y /= self.sampling ** 2

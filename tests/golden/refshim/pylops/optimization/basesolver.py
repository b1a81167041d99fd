import time


class Solver:
    def __init__(self, Op, callbacks=None):
        self.Op = Op
        self.callbacks = callbacks
        self.tstart = time.time()

    def callback(self, x, *args, **kwargs):
        pass

    def _print_solver(self, *a, **k):
        pass

    def _print_finalize(self, *a, **k):
        pass

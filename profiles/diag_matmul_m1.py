"""Where does the M=1 MPIMatrixMult apply (32768^2 bf16, 1 GPU) spend its time: kernel vs host path?"""
import cProfile
import io
import json
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pylops_mpi_b200 as pm  # noqa: E402
from pylops_mpi_b200 import _lib as L  # noqa: E402


def gpu_ms(fn, k=20, w=5):
    for _ in range(w):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(k):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / k


def host_ms(fn, k=20):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(k):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / k * 1e3


out = {}
n = 32768
A = (torch.randn(n, n, device="cuda", dtype=torch.float32) / 181).to(torch.bfloat16)
x = torch.randn(n, device="cuda")
y = torch.empty(n, device="cuda")
for op, name in ((L.OP_N, "N"), (L.OP_T, "T")):
    def k():
        L.check(L.lib.b2_gemv(L.ctx(), A.data_ptr(), n, n, n, x.data_ptr(), y.data_ptr(), op, L.BF16, L.F32, L.stream()))
    ms = gpu_ms(k)
    out[f"raw_gemv_bf16_{name}"] = {"ms": ms, "GB/s": 2 * n * n / ms / 1e6}
for rep in (False, True):
    Op = pm.MPIMatrixMult(A, 1, kind="summa", dtype="bfloat16", replicate=rep) if rep else \
        pm.MPIMatrixMult(A, 1, kind="summa", dtype="bfloat16")
    xd = pm.DistributedArray(global_shape=n, dtype=np.float32)
    xd.local_array.normal_()
    tag = "replicated" if rep else "summa"
    out[f"{tag}_matvec"] = {"gpu_ms": gpu_ms(lambda: Op.matvec(xd)), "host_enqueue_ms": host_ms(lambda: Op.matvec(xd))}
    yd = Op.matvec(xd)
    out[f"{tag}_rmatvec"] = {"gpu_ms": gpu_ms(lambda: Op.rmatvec(yd)), "host_enqueue_ms": host_ms(lambda: Op.rmatvec(yd))}
    if not rep:
        pr = cProfile.Profile()
        pr.enable()
        for _ in range(50):
            Op.matvec(xd)
        pr.disable()
        s = io.StringIO()
        pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(18)
        out["cprofile_matvec_50"] = s.getvalue().splitlines()[4:40]
    del Op
print(json.dumps(out, indent=1))

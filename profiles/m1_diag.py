"""Where the time of the M = 1 ("32768-vec") MPIMatrixMult apply goes at N > 1: the per-rank GEMV alone, the small
all-gather of x alone, the operator call, and the host enqueue cost of the operator call.

    torchrun --nproc-per-node 2 profiles/m1_diag.py
"""
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
from pylops_mpi_b200 import _lib  # noqa: E402
from pylops_mpi_b200.Distributed import allgatherv  # noqa: E402
from pylops_mpi_b200.basicoperators.MatrixMult import tile_product  # noqa: E402


def ev_time(fn, n=30, warm=10):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    pm.COMM_WORLD.Barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    enq = (time.perf_counter() - t) / n * 1e6
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3, enq


def main():
    comm = pm.COMM_WORLD
    P, rank = comm.Get_size(), comm.Get_rank()
    Ng = Kg = 32768
    A = (torch.randn(Ng // P, Kg, device="cuda") / 181).to(torch.bfloat16)
    Sop = pm.MPIMatrixMult(A, 1, kind="summa", dtype="bfloat16", grid=(P, 1), replicate=True)
    xs = pm.DistributedArray(global_shape=Kg, local_shapes=[Kg // P] * P, dtype=np.float32)
    xs.local_array.normal_()
    xfull = torch.randn(Kg, 1, device="cuda")
    yloc = torch.empty(Ng // P, 1, device="cuda")
    out = {"n_gpus": P}
    out["gemv_alone_us"], out["gemv_enqueue_us"] = ev_time(lambda: tile_product(A, xfull, yloc, _lib.OP_N, False))
    out["gemv_GBps"] = 2.0 * A.numel() / out["gemv_alone_us"] / 1e3
    if P > 1:
        xl = xs.local_array
        out["allgather_alone_us"], out["allgather_enqueue_us"] = ev_time(
            lambda: allgatherv(comm, xl, [Kg // P] * P))
    out["matvec_us"], out["matvec_enqueue_us"] = ev_time(lambda: Sop.matvec(xs))
    out["rmatvec_us"], out["rmatvec_enqueue_us"] = ev_time(lambda: Sop.rmatvec(xs))
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(300):
        Sop.matvec(xs)
    pr.disable()
    torch.cuda.synchronize()
    comm.Barrier()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(16)
    if rank == 0:
        print(json.dumps(out))
        print(s.getvalue())


if __name__ == "__main__":
    main()

"""Host (Python) cost of ONE ``MPIFirstDerivative.matvec`` call: the kernel is made tiny (4096 x 512 rows per
GPU) so the wall time of a long un-synchronised loop IS the enqueue cost.  At N = 8 the headline step is 81 us of
kernel, so the enqueue must stay well below that (round 2, 8-GPU run: 88 us before the trims -> host-bound).

    python profiles/host_enqueue.py                    # N = 1 (plain kernel)
    torchrun --nproc-per-node 2 profiles/host_enqueue.py   # peer-halo path
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


def main():
    comm = pm.COMM_WORLD
    P, rank = comm.Get_size(), comm.Get_rank()
    dims = (4096 * P, 512)
    op = pm.MPIFirstDerivative(dims, kind="centered", order=3, dtype=np.float32)
    x = pm.DistributedArray(global_shape=int(np.prod(dims)), dtype=np.float32)
    x.local_array.normal_()
    for _ in range(200):
        y = op.matvec(x)
    torch.cuda.synchronize()
    comm.Barrier()
    n = 3000
    t0 = time.perf_counter()
    for _ in range(n):
        y = op.matvec(x)
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    t2 = time.perf_counter()
    out = {"n_gpus": P, "rank": rank, "enqueue_us_per_matvec": (t1 - t0) / n * 1e6,
           "drain_us": (t2 - t1) * 1e6}
    comm.Barrier()
    pr = cProfile.Profile()
    pr.enable()
    for _ in range(1000):
        y = op.matvec(x)
    pr.disable()
    torch.cuda.synchronize()
    comm.Barrier()
    s = io.StringIO()
    pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(18)
    if rank == 0:
        print(json.dumps(out))
        print(s.getvalue())
    del y


if __name__ == "__main__":
    main()

"""Cost of the one-off CUDA-graph capture inside CGLS.run (config 3: 4096^2 float32 block per GPU): three
consecutive 50-iteration solves, wall time and the capture breakdown of each (B2_CGLS_CAPTURE_MODE selects torch's
capture error mode)."""
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pylops_mpi_b200 as pm  # noqa: E402

n = 4096
A = torch.randn(n, n, device="cuda") / 128 + 2 * torch.eye(n, device="cuda")
Op = pm.MPIBlockDiag([pm.MatrixMult(A)])
xt = pm.DistributedArray(global_shape=n * pm.COMM_WORLD.Get_size(), dtype=np.float32)
xt.local_array.normal_()
y = Op @ xt
x0 = xt.zeros_like() if hasattr(xt, "zeros_like") else None
rows = []
for i in range(4):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    s = pm.CGLS(Op)
    s.solve(y, x0=x0, niter=50, tol=0.0)
    torch.cuda.synchronize()
    rows.append({"solve": i, "ms": (time.perf_counter() - t0) * 1e3, "capture_ms": getattr(s, "graph_capture_ms", None),
                 "breakdown": getattr(s, "graph_capture_breakdown_ms", None), "replays": s.graph_replays,
                 "error": s.graph_error})
    del s
if pm.COMM_WORLD.Get_rank() == 0:
    print(json.dumps({"mode": os.environ.get("B2_CGLS_CAPTURE_MODE", "thread_local"), "solves": rows}))

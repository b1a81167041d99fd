"""bf16 GEMV (op N) on 32768-column row panels of 32768 / 16384 / 8192 / 4096 rows -- the per-GPU share of the
"32768-vec" MPIMatrixMult at N = 1, 2, 4, 8 -- and the float32 4096^2 block of CGLS config 3 (L2 resident).
Run once with B2_GEMV_SPLIT=1 (default) and once with B2_GEMV_SPLIT=0 (one warp per row)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pylops_mpi_b200 import _lib as L  # noqa: E402

out = {"B2_GEMV_SPLIT": os.environ.get("B2_GEMV_SPLIT", "1"), "cases": []}
n = 32768
x = torch.randn(n, device="cuda")
for m in (32768, 16384, 8192, 4096):
    A = (torch.randn(m, n, device="cuda") / 181).to(torch.bfloat16)
    y = torch.empty(m, device="cuda")

    def fn():
        L.check(L.lib.b2_gemv(L.ctx(), A.data_ptr(), n, m, n, x.data_ptr(), y.data_ptr(), 0, L.BF16, L.F32, L.stream()))
    for _ in range(10):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        fn()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    ref = A[:64].double() @ x.double()
    err = float((y[:64].double() - ref).abs().max())
    out["cases"].append({"rows": m, "us": us, "GB/s": 2.0 * m * n / us / 1e3, "max_abs_err_64_rows": err})
    del A, y
for m, nn in ((4096, 4096), (32768, 8192)):
    A = torch.randn(m, nn, device="cuda")
    xx = torch.randn(nn, device="cuda")
    y = torch.empty(m, device="cuda")

    def fn2():
        L.check(L.lib.b2_gemv(L.ctx(), A.data_ptr(), nn, m, nn, xx.data_ptr(), y.data_ptr(), 0, L.F32, L.F32, L.stream()))
    for _ in range(10):
        fn2()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30):
        fn2()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / 30 * 1e3
    out["cases"].append({"f32": [m, nn], "us": us, "GB/s": 4.0 * m * nn / us / 1e3})
print(json.dumps(out))

"""Time the fused ISTA / FISTA model-update kernel alone (CUDA events, inputs >> L2)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pylops_mpi_b200 as pm  # noqa: E402,F401
from pylops_mpi_b200 import _lib as L  # noqa: E402
from pylops_mpi_b200.optimization.cls_sparsity import _sparse_update  # noqa: E402

PEAK = json.load(open(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json"))).get("hbm_gbs", 6578.0) \
    if os.path.exists(os.path.join(os.path.dirname(__file__), "..", "MEASURED_PEAKS.json")) else 6578.0


def timeit(fn, k=10, w=3):
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


out = {}
sums = torch.zeros(4, dtype=torch.float64, device="cuda")
for name, dt, n in (("f32", torch.float32, 1 << 28), ("f64", torch.float64, 1 << 27), ("c64", torch.complex64, 1 << 27)):
    x = torch.randn(n, dtype=dt, device="cuda")
    g = torch.randn(n, dtype=dt, device="cuda")
    z = torch.randn(n, dtype=dt, device="cuda")
    es = x.element_size()
    for kind, kn in ((L.THRESH_SOFT, "soft"), (L.THRESH_HARD, "hard"), (L.THRESH_HALF, "half")):
        if dt.is_complex and kn == "half":
            continue
        ms = timeit(lambda: _sparse_update(x, g, 1e-3, x, 1e-4, kind, x, None, 0.0, sums))
        out[f"ista_{name}_{kn}"] = {"us": ms * 1e3, "GB/s": 3 * es * n / ms / 1e6, "frac_hbm": 3 * es * n / ms / 1e6 / PEAK}
        ms = timeit(lambda: _sparse_update(z, g, 1e-3, x, 1e-4, kind, x, z, 0.3, sums))
        out[f"fista_{name}_{kn}"] = {"us": ms * 1e3, "GB/s": 5 * es * n / ms / 1e6, "frac_hbm": 5 * es * n / ms / 1e6 / PEAK}
    del x, g, z
print(json.dumps(out, indent=1))

"""Time the stencil kernel variants (B2_STENCIL_VARIANT) on the headline shape; one process per
variant because the choice is latched at first use.  Run on the GPU box."""
import json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CODE = r'''
import sys, torch
sys.path.insert(0, %r)
from pylops_mpi_b200 import _lib as L
res = {}
for name, dt, code, kind, order, adj, rows in (("c3_f32", torch.float32, L.F32, 2, 3, 0, 32768), ("c5_f32", torch.float32, L.F32, 2, 5, 0, 32768),
                                         ("c3_f64", torch.float64, L.F64, 2, 3, 0, 16384), ("c3adj_f32", torch.float32, L.F32, 2, 3, 1, 32768)):
    x = torch.randn(rows, 8192, device="cuda", dtype=dt); y = torch.empty_like(x)
    def k():
        L.check(L.lib.b2_first_derivative(L.ctx(), x.data_ptr(), y.data_ptr(), None, 0, None, 0, rows, 8192, 0, rows, kind, order, 0, 1.0, adj, code, L.stream()))
    for _ in range(5): k()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(30): k()
    e1.record(); torch.cuda.synchronize()
    res[name] = 2 * x.element_size() * x.numel() * 30 / (e0.elapsed_time(e1) * 1e-3) / 1e9
print(res)
''' % ROOT
out = {}
for v in range(12):
    r = subprocess.run([sys.executable, "-c", CODE], env=dict(os.environ, B2_STENCIL_VARIANT=str(v)), capture_output=True, text=True)
    out[v] = r.stdout.strip() or r.stderr[-300:]
    print(v, out[v], flush=True)

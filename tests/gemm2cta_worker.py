"""Checks the cta_group::2 tensor-core kernel (B2_GEMM_2CTA=1 must be set by the launcher)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pylops_mpi_b200 import _lib as L  # noqa: E402

VARIANT = os.environ.get("B2_GEMM_2CTA", "default")
for (m, n, k) in [(256, 256, 64), (256, 256, 512), (512, 512, 256), (1024, 1024, 1024), (384, 264, 72),
                  (200, 40, 1000), (2048, 768, 320), (129, 256, 64)]:
    for op in (0, 1):
        torch.manual_seed(m + n + k + op)
        A = (torch.randn((m, k) if op == 0 else (k, m), device="cuda") / 8).to(torch.bfloat16)
        if op == 1 and m % 8:
            continue
        B = (torch.randn(k, n, device="cuda") / 8).to(torch.bfloat16)
        C = torch.full((m, n), 7.0, device="cuda")
        L.check(L.lib.b2_gemm_bf16(L.ctx(), A.data_ptr(), A.shape[1], B.data_ptr(), n, C.data_ptr(), n, m, n, k, op, 0,
                                   L.stream()), "gemm 2cta")
        torch.cuda.synchronize()
        A64 = A.double() if op == 0 else A.double().T
        ref = A64 @ B.double()
        bound = (A64.abs() @ B.double().abs()) * (k * 6e-8) + 1e-6
        err = (C.double() - ref).abs()
        assert bool((err <= bound).all()), f"2cta mismatch {m},{n},{k},{op}: {err.max().item()}"
        L.check(L.lib.b2_gemm_bf16(L.ctx(), A.data_ptr(), A.shape[1], B.data_ptr(), n, C.data_ptr(), n, m, n, k, op, 1,
                                   L.stream()), "gemm 2cta acc")
        assert bool(((C.double() - 2 * ref).abs() <= 2 * bound + 1e-5).all())
# speed
m = n = k = 8192
A = (torch.randn(m, k, device="cuda") / 90).to(torch.bfloat16)
B = (torch.randn(k, n, device="cuda") / 90).to(torch.bfloat16)
C = torch.empty(m, n, device="cuda")
f = lambda: L.check(L.lib.b2_gemm_bf16(L.ctx(), A.data_ptr(), k, B.data_ptr(), n, C.data_ptr(), n, m, n, k, 0, 0, L.stream()))  # noqa: E731
for _ in range(3):
    f()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    f()
e1.record()
torch.cuda.synchronize()
print("GEMM2CTA_OK variant", VARIANT, "TF/s", 2.0 * m * n * k * 10 / (e0.elapsed_time(e1) * 1e-3) / 1e12)

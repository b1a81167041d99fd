"""Round-2 check + timing of the tcgen05 (bf16x3) MPIFredholm1 product against a float64 / complex128 torch
reference (test infrastructure; run on the GPU box):   python profiles/fredholm_tc_check.py [--time]
B2_FREDHOLM_BK=64|32 selects the operand ring (read once per process by the library)."""
import ctypes as C
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from pylops_mpi_b200 import _lib as L  # noqa: E402


def run(nsl, nx, ny, nz, dtype, adjoint, seed=0, scaled=False):
    g = torch.Generator(device="cuda").manual_seed(seed)
    cx = dtype is torch.complex64
    G = torch.randn(nsl, nx, ny, device="cuda", dtype=dtype, generator=g)
    kin, kout = (nx, ny) if adjoint else (ny, nx)
    x = torch.randn(nsl, kin, nz, device="cuda", dtype=dtype, generator=g)
    if scaled:     # wide dynamic range: per-slice scales of G, per-column scales of x, a few huge / tiny entries
        G = G * (10.0 ** torch.randint(-15, 16, (nsl, 1, 1), device="cuda", generator=g).float())
        x = x * (10.0 ** torch.randint(-8, 9, (nsl, 1, nz), device="cuda", generator=g).float())
        G.view(-1)[::97] *= 1e-6
        x.view(-1)[::89] *= 1e-5
    y = torch.full((nsl, kout, nz), float("nan"), device="cuda", dtype=dtype)
    h = C.c_void_p()
    L.check(L.lib.b2_fredholm_plan_create(L.ctx(), G.data_ptr(), nsl, nx, ny, nz, L.code(dtype), C.byref(h)), "plan")
    L.check(L.lib.b2_fredholm_apply(h, x.data_ptr(), y.data_ptr(), None, 0, int(adjoint), L.stream()), "apply")
    torch.cuda.synchronize()
    wide = torch.complex128 if cx else torch.float64
    Gw = G.to(wide)
    ref = torch.matmul(Gw.conj().transpose(1, 2) if adjoint else Gw, x.to(wide))
    # worst (slice, column): every column of every slice must be accurate relative to ITS OWN size
    colerr = ((y.to(wide) - ref).abs().amax(dim=1) / ref.abs().amax(dim=1).clamp_min(1e-300))
    err = colerr.max().item()
    nrm = ((y.to(wide) - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-300)).max().item()
    # SIMT kernel on the same inputs
    ys = torch.empty_like(y)
    L.check(L.lib.b2_batched_gemm(L.ctx(), G.data_ptr(), x.data_ptr(), ys.data_ptr(), nsl, nx, ny, nz, int(adjoint),
                                  L.code(dtype), L.stream()), "simt")
    torch.cuda.synchronize()
    nrm_simt = ((ys.to(wide) - ref).norm(dim=1) / ref.norm(dim=1).clamp_min(1e-300)).max().item()
    L.lib.b2_fredholm_plan_destroy(h)
    return err, nrm, nrm_simt


def main():
    mode, bk = os.environ.get("B2_FREDHOLM_MODE", "h2"), os.environ.get("B2_FREDHOLM_BK", "32")
    out = {"mode": mode, "bk": bk, "cases": []}
    bad = 0
    shapes = [(64, 256, 256, 64), (3, 128, 128, 64), (5, 100, 70, 9), (21, 4, 6, 5), (2, 300, 130, 70), (4, 17, 33, 1),
              (2, 129, 257, 65), (1, 512, 64, 128)]
    if "--time-only" in sys.argv:
        shapes = []
    for shp in shapes:
        for dtype in (torch.complex64, torch.float32):
            for adj in (False, True):
                for scaled in (False, True):
                    err, nrm, nrm_simt = run(*shp, dtype, adj, scaled=scaled)
                    ok = nrm < 3e-6 and err < 1e-5
                    bad += (not ok)
                    out["cases"].append({"shape": shp, "dtype": str(dtype), "adjoint": adj, "scaled": scaled,
                                         "worst_column_max_err": err, "worst_column_normwise_err": nrm,
                                         "worst_column_normwise_err_simt_fp32": nrm_simt, "ok": ok})
                    print(shp, dtype, adj, "scaled" if scaled else "plain",
                          f"maxerr {err:.2e} normwise {nrm:.2e} (simt {nrm_simt:.2e})", "OK" if ok else "FAIL", flush=True)
    if "--time" in sys.argv:
        nsl, nx, ny, nz = 64, 256, 256, 64
        G = torch.randn(nsl, nx, ny, device="cuda", dtype=torch.complex64)
        x = torch.randn(nsl, ny, nz, device="cuda", dtype=torch.complex64)
        y = torch.empty(nsl, nx, nz, device="cuda", dtype=torch.complex64)
        h = C.c_void_p()
        L.check(L.lib.b2_fredholm_plan_create(L.ctx(), G.data_ptr(), nsl, nx, ny, nz, L.C64, C.byref(h)), "plan")
        st = L.stream()

        def tc():
            L.lib.b2_fredholm_apply(h, x.data_ptr(), y.data_ptr(), None, 0, 0, st)

        def simt():
            L.lib.b2_batched_gemm(L.ctx(), G.data_ptr(), x.data_ptr(), y.data_ptr(), nsl, nx, ny, nz, 0, L.C64, st)

        def tc_pack():
            L.lib.b2_fredholm_apply_parts(h, x.data_ptr(), y.data_ptr(), 0, 1, st)

        def tc_product():
            L.lib.b2_fredholm_apply_parts(h, x.data_ptr(), y.data_ptr(), 0, 2, st)

        for name, fn in (("tc_pack+product", tc), ("tc_pack_only", tc_pack), ("tc_product_only", tc_product), ("simt", simt)):
            for _ in range(10):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(5 if "--time-only" in sys.argv else 50):
                fn()
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / (5 if "--time-only" in sys.argv else 50) * 1e3
            out[name + "_us"] = us
            print(name, f"{us:.1f} us  ({8.0 * nsl * nx * ny * nz / us / 1e6:.1f} TF/s complex-equivalent)", flush=True)
    out["failed"] = bad
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", f"fredholm_tc_check_{mode}_bk{bk}.json"), "w") as f:
        json.dump(out, f, indent=1)
    print("FAILED" if bad else "ALL OK", bad)
    return 1 if bad else 0


if __name__ == "__main__":
    sys.exit(main())

#!/usr/bin/env python
"""Benchmark of the pylops-mpi hot path on B200 (contract: see the task brief).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--no-extras]

Headline (BASELINE.json metric "... GB/s (FirstDerivative)"): one *step* is one
``MPIFirstDerivative.matvec`` (centered, order 3, float32) over a row-block
partitioned array with 32768 x 8192 elements PER GPU (1 GiB in + 1 GiB out per
GPU, >> the 126 MB L2; weak scaling: global rows = 32768 * N), inputs resident
in HBM.  ``value`` = algorithmic bytes (2 * 4 B per element, all ranks) / time.
``e2e`` = the same operator through the host-buffer plugin entry
(``b2_first_derivative_host``: pinned host arrays in, pinned host arrays out,
H2D + kernel + D2H inside the timed region).  ``extra`` carries the other
BASELINE configs (reductions, BlockDiag GEMV / cgls, MatrixMult GF/s,
Fredholm1) measured in the same run.

``--impl reference`` times the reference's CPU algorithm for the same operator
(the NumPy restatement in oracle/, one OS process per host core, each applying
the per-rank stencil code to its own row block) on a bounded sample.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "oracle")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

# rank 0 must print exactly ONE JSON line on stdout: keep NCCL's banner / debug text off it
os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")

ROWS_PER_GPU = 32768
NCOLS = 8192
METRIC = "MPIFirstDerivative matvec GB/s (algorithmic bytes, centered-3 float32)"
HBM_FALLBACK = 6650.0


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return json.load(f), "measured"
    except Exception:
        return {"hbm_gbs": HBM_FALLBACK, "bf16_tflops": 1590.0}, "fallback"


# --------------------------------------------------------------------------
# clocks sampler (nvidia-smi during the timed region)
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int = 0):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], None, set()
        for ln in self.lines:
            parts = [p.strip() for p in ln.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx = float(parts[1])
            except ValueError:
                continue
            for name, val in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons),
                "samples": len(sm)}


# --------------------------------------------------------------------------
# CPU reference arm (oracle port, one process per core)
# --------------------------------------------------------------------------
def _cpu_rank_main(conn, rows, ncols, seed):
    """one simulated CPU rank: owns its row block (created ONCE, outside the timed passes) and applies the
    reference's per-rank stencil code to it every time the parent says go"""
    import pylops_mpi_oracle as o
    x = np.random.default_rng(seed).standard_normal((rows, ncols), dtype=np.float32)
    flat = [x.ravel()]
    conn.send("ready")
    while True:
        msg = conn.recv()
        if msg is None:
            break
        t0 = time.perf_counter()
        y = o.first_derivative(flat, (rows, ncols), 1.0, "centered", False, 3, False, dtype=np.float32)
        conn.send((time.perf_counter() - t0, float(y[0][ncols + 1])))


class CpuRanks:
    """`cores` OS processes, each a simulated rank of the reference's NumPy path"""

    def __init__(self, cores: int, rows_per_proc: int, ncols: int):
        import multiprocessing as mp
        ctx = mp.get_context("spawn")
        self.cores = cores
        self.conns, self.procs = [], []
        for i in range(cores):
            a, b = ctx.Pipe()
            p = ctx.Process(target=_cpu_rank_main, args=(b, rows_per_proc, ncols, 42 + i), daemon=True)
            p.start()
            self.conns.append(a)
            self.procs.append(p)
        for c in self.conns:
            c.recv()

    def step(self) -> float:
        """one pass: every rank applies the stencil to its block concurrently; wall-clock seconds"""
        t0 = time.perf_counter()
        for c in self.conns:
            c.send("go")
        for c in self.conns:
            c.recv()
        return time.perf_counter() - t0

    def close(self):
        for c in self.conns:
            c.send(None)
        for p in self.procs:
            p.join(timeout=10)

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False


def cpu_reference_pass(pool, cores: int, rows_per_proc: int, ncols: int, reps: int = 1):
    """one 'step' of the CPU arm: every process applies the per-rank reference stencil
    (oracle.first_derivative -> FirstDerivative.py:201-219 incl. its temporaries) to its block"""
    return pool.step()


def run_reference_arm(args):
    import multiprocessing as mp
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = min(os.cpu_count() or 1, 64)
    rows_per_proc = 2048           # 64 MiB float32 per process per pass
    with CpuRanks(cores, rows_per_proc, NCOLS) as pool:
        for _ in range(max(1, args.warmup // 2)):
            cpu_reference_pass(pool, cores, rows_per_proc, NCOLS)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            cpu_reference_pass(pool, cores, rows_per_proc, NCOLS)
        dt = time.perf_counter() - t0
    nbytes = 2 * 4 * rows_per_proc * NCOLS * cores * args.steps
    val = nbytes / dt / 1e9
    sample = f"{cores} processes x ({rows_per_proc} x {NCOLS}) float32 rows per step (bounded sample of the {ROWS_PER_GPU} x {NCOLS} per-GPU block)"
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": "GB/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(args.gpus),
            "cpu_baseline": {"value": val, "unit": "GB/s", "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": val, "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def workload_config(n):
    return {"workload": "MPIFirstDerivative matvec, kind=centered order=3 edge=False sampling=1, float32, "
                        f"dims=({ROWS_PER_GPU}*N, {NCOLS}) row-block partition, N={n}",
            "rows_per_gpu": ROWS_PER_GPU, "ncols": NCOLS, "global_rows": ROWS_PER_GPU * n,
            "l2": "per-GPU input 1 GiB >> 126 MB L2 (no flush needed)", "parallelism": f"rows{n}"}


# --------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------
def time_loop(fn, steps, warmup, comm=None):
    import torch
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    if comm is not None:
        comm.Barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    t_cpu = time.perf_counter()
    for _ in range(steps):
        fn()
    time_loop.last_enqueue_ms = (time.perf_counter() - t_cpu) * 1e3 / max(steps, 1)
    e1.record()
    torch.cuda.synchronize()
    if comm is not None:
        comm.Barrier()
    ms = e0.elapsed_time(e1)
    if comm is not None and comm.Get_size() > 1:
        ms = comm.allreduce(ms, "max")
    return ms


def run_gpu_arm(args):
    import torch
    import pylops_mpi_b200 as pm
    from pylops_mpi_b200 import _lib as L

    comm = pm.get_comm_world()
    rank, size = comm.Get_rank(), comm.Get_size()
    if size != args.gpus:
        if rank == 0:
            print(f"warning: --gpus {args.gpus} but world size {size}", file=sys.stderr)
    peaks, peak_kind = load_peaks()
    dev = torch.cuda.current_device()
    N = ROWS_PER_GPU * size
    dims = (N, NCOLS)
    nloc = ROWS_PER_GPU
    elem_loc = nloc * NCOLS
    bytes_loc = 2 * 4 * elem_loc

    # ---- inputs resident in HBM --------------------------------------------------
    g = torch.Generator(device="cuda").manual_seed(42 + rank)
    x = pm.DistributedArray(global_shape=N * NCOLS, dtype=np.float32)
    assert x.local_shape == (elem_loc,)
    x.local_array.normal_(generator=g)
    Fop = pm.MPIFirstDerivative(dims, kind="centered", order=3, dtype=np.float32)
    holder = {}

    def step():
        holder["y"] = Fop.matvec(x)

    sampler = ClockSampler(dev)
    sampler.start()                      # samples cover warm-up, the timed region and the kernel-only loop
    for _ in range(args.warmup):
        step()
    ms = time_loop(step, args.steps, 0, comm)
    enqueue_ms = time_loop.last_enqueue_ms     # host time to enqueue one step (GPU-bound if << ms_per_step)
    value = bytes_loc * size * args.steps / (ms * 1e-3) / 1e9

    # ---- roofline of the dominant kernel: live CUDA-event timing of the kernel alone --------
    xl = x.local_array
    yl = torch.empty_like(xl)
    st = L.stream()

    def kern():
        L.check(L.lib.b2_first_derivative(L.ctx(), xl.data_ptr(), yl.data_ptr(), None, 0, None, 0, nloc, NCOLS,
                                          0, nloc, L.FD_CENTERED, 3, 0, 1.0, 0, L.F32, st))
    kms = time_loop(kern, args.steps, max(3, args.warmup), None) / args.steps
    achieved = bytes_loc / (kms * 1e-3) / 1e9
    # the timed region is only a few ms: keep the same kernel running ~0.7 s so that the 100 ms
    # nvidia-smi sampler sees clocks / throttle reasons under this exact load
    t_load = time.perf_counter()
    while time.perf_counter() - t_load < 0.7:
        for _ in range(20):
            kern()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    traffic = None
    try:
        with open(os.path.join(ROOT, "profiles", "traffic.json")) as f:
            traffic = json.load(f).get("stencil_vec_kernel_f32_centered3_bytes_per_launch")
    except Exception:
        pass
    roofline = {"bound": "hbm", "kernel": "stencil_vec_kernel<float, taps{-1,+1}>", "achieved": achieved,
                "peak": peaks["hbm_gbs"], "peak_kind": f"{peak_kind} copy bandwidth (burst)", "unit": "GB/s",
                "frac": achieved / peaks["hbm_gbs"], "traffic": traffic,
                "algorithmic_bytes_per_launch": bytes_loc, "us_per_launch": kms * 1e3}

    # ---- e2e: host buffers through the plugin entry (H2D + kernel + D2H in the timed region) --
    # each rank owns rows [rank*nloc, (rank+1)*nloc) of the replicated host array; only its block
    # (+2 halo rows each side) is materialised, addressed through a virtual global base pointer
    lo = 2 if rank > 0 else 0
    hi = 2 if rank < size - 1 else 0
    xh = torch.empty((nloc + lo + hi, NCOLS), dtype=torch.float32).pin_memory()
    xh.normal_()
    yh = torch.empty((nloc, NCOLS), dtype=torch.float32).pin_memory()
    row_bytes = NCOLS * 4
    x_base = xh.data_ptr() - (rank * nloc - lo) * row_bytes
    y_base = yh.data_ptr() - (rank * nloc) * row_bytes

    def e2e_step():
        L.check(L.lib.b2_first_derivative_host(L.ctx(), x_base, y_base, N, NCOLS, rank * nloc, (rank + 1) * nloc,
                                               L.FD_CENTERED, 3, 0, 1.0, 0, L.F32), "b2_first_derivative_host")
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        e2e_step()
    comm.Barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    if size > 1:
        e2e_s = comm.allreduce(e2e_s, "max")
    e2e_val = bytes_loc * size * e2e_steps / e2e_s / 1e9
    e2e = {"value": e2e_val, "unit": "GB/s", "h2d_bytes_per_step": (nloc + lo + hi) * row_bytes * size,
           "d2h_bytes_per_step": nloc * row_bytes * size, "steps": e2e_steps,
           "api": "b2_first_derivative_host (pinned host in/out, 3-stream chunk pipeline)"}
    # spot-check the e2e result against the device path (same data -> same numbers)
    xchk = torch.as_tensor(xh[lo:lo + 64]).cuda()
    del xh, yh, xchk

    extra = {}
    if not args.no_extras:
        try:
            extra = run_extras(pm, L, comm, peaks, args)
        except Exception as exc:  # extras must never kill the headline line
            extra = {"error": repr(exc)}

    cpu_baseline = None
    if rank == 0 and size == 1 and not args.no_cpu:
        import multiprocessing as mp
        cores = min(os.cpu_count() or 1, 64)
        rows_per_proc = 2048
        with CpuRanks(cores, rows_per_proc, NCOLS) as pool:
            cpu_reference_pass(pool, cores, rows_per_proc, NCOLS)
            reps = 0
            t0 = time.perf_counter()
            while time.perf_counter() - t0 < 10.0 and reps < 50:
                cpu_reference_pass(pool, cores, rows_per_proc, NCOLS)
                reps += 1
            dt = time.perf_counter() - t0
        cpu_baseline = {"value": 2 * 4 * rows_per_proc * NCOLS * cores * reps / dt / 1e9, "unit": "GB/s",
                        "cores": cores, "kind": "port",
                        "sample": f"{reps} passes of {cores} processes x ({rows_per_proc} x {NCOLS}) float32 "
                                  "(oracle restatement of FirstDerivative.py:201-219 per rank)"}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "GB/s", "n_gpus": size, "steps": args.steps,
                "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
                "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": workload_config(size), "roofline": roofline, "cpu_baseline": cpu_baseline,
                "e2e": e2e, "gpu_launches": args.steps * (1 if size == 1 else 3), "clocks": clocks,
                "host_enqueue_ms_per_step": enqueue_ms, "extra": extra}
        print(json.dumps(line))


def run_extras(pm, L, comm, peaks, args):
    """other BASELINE configs, each timed with CUDA events after warm-up (short loops)"""
    import torch
    rank, size = comm.Get_rank(), comm.Get_size()
    out = {}
    K, W = 10, 3
    hbm = peaks["hbm_gbs"]

    def gbs(nbytes, ms_total, k=K):
        return nbytes * k / (ms_total * 1e-3) / 1e9

    # --- FirstDerivative variants (per-GPU kernel, device resident) -------------------
    nloc, ncols = ROWS_PER_GPU, NCOLS
    for name, dt, code, kind, order, adj in (("fd_centered3_adj_f32", torch.float32, L.F32, L.FD_CENTERED, 3, 1),
                                             ("fd_centered5_f32", torch.float32, L.F32, L.FD_CENTERED, 5, 0),
                                             ("fd_forward_f32", torch.float32, L.F32, L.FD_FORWARD, 3, 0),
                                             ("fd_centered3_f64", torch.float64, L.F64, L.FD_CENTERED, 3, 0)):
        rows = nloc if dt is torch.float32 else nloc // 2
        xl = torch.randn(rows, ncols, device="cuda", dtype=dt)
        yl = torch.empty_like(xl)

        def kern():
            L.check(L.lib.b2_first_derivative(L.ctx(), xl.data_ptr(), yl.data_ptr(), None, 0, None, 0, rows, ncols,
                                              0, rows, kind, order, 0, 1.0, adj, code, L.stream()))
        ms = time_loop(kern, K, W)
        v = gbs(2 * xl.element_size() * xl.numel(), ms)
        out[name] = {"GB/s": v, "frac_hbm": v / hbm}
        del xl, yl

    # --- config 2: local reductions (dot 8 B/elem, norm 4 B/elem) ------------------------
    n = 1 << 28
    a = pm.DistributedArray(global_shape=n * size, dtype=np.float32)
    b = pm.DistributedArray(global_shape=n * size, dtype=np.float32)
    a.local_array.normal_()
    b.local_array.normal_()
    ms = time_loop(lambda: a._dot_device(b), K, W, comm)
    out["dot_f32_2^28_per_gpu"] = {"GB/s": gbs(8 * n * size, ms), "frac_hbm": gbs(8 * n, ms) / hbm, "us": ms / K * 1e3}
    ms = time_loop(lambda: a._norm_device(2), K, W, comm)
    out["norm2_f32_2^28_per_gpu"] = {"GB/s": gbs(4 * n * size, ms), "frac_hbm": gbs(4 * n, ms) / hbm, "us": ms / K * 1e3}
    # --- "next" row: fused ISTA / FISTA model update (ISTA: 2 reads + 1 write, FISTA: 3 reads + 2 writes per elem) ---
    from pylops_mpi_b200.optimization.cls_sparsity import _sparse_update
    sums = torch.zeros(4, dtype=torch.float64, device="cuda")
    xa, ga = a.local_array, b.local_array
    ms = time_loop(lambda: _sparse_update(xa, ga, 1e-3, xa, 1e-4, L.THRESH_SOFT, xa, None, 0.0, sums), K, W)
    out["ista_update_f32_2^28"] = {"GB/s": gbs(12 * n, ms), "frac_hbm": gbs(12 * n, ms) / hbm, "us": ms / K * 1e3}
    za = torch.randn_like(xa)
    ms = time_loop(lambda: _sparse_update(za, ga, 1e-3, xa, 1e-4, L.THRESH_SOFT, xa, za, 0.3, sums), K, W)
    out["fista_update_f32_2^28"] = {"GB/s": gbs(20 * n, ms), "frac_hbm": gbs(20 * n, ms) / hbm, "us": ms / K * 1e3}
    del za
    small = pm.DistributedArray(global_shape=10000 * size, dtype=np.float32)
    small.local_array.normal_()
    ms = time_loop(lambda: small.dot(small), 50, 5, comm)
    out["dot_f32_1e4_host_result_us"] = ms / 50 * 1e3
    if size > 1:
        from pylops_mpi_b200.Distributed import allreduce_
        sweep = {}
        big = torch.zeros(10 ** 9, dtype=torch.float32, device="cuda")      # 4 GB: config 2 sweeps 1e4 .. 1e9 elements
        for ne in (10 ** 4, 10 ** 5, 10 ** 6, 10 ** 7, 10 ** 8, 10 ** 9):
            buf = big[:ne]
            ms = time_loop(lambda: allreduce_(comm, buf), 10, 3, comm)
            us = ms / 10 * 1e3
            alg = 4 * ne / (us * 1e-6) / 1e9
            sweep[str(ne)] = {"us": us, "algbw_GB/s": alg, "busbw_GB/s": alg * 2 * (size - 1) / size,
                              "path": "peer-memory one-shot" if 4 * ne <= 64 * 1024 and comm.peer_vec is not None else "nccl"}
        out["allreduce_f32_sweep"] = sweep
        del big
    del a, b

    # --- config 3: BlockDiag of one 4096^2 f32 block per GPU + cgls ------------------------
    nb = 4096
    A = torch.randn(nb, nb, device="cuda", generator=torch.Generator(device="cuda").manual_seed(100 + rank)) / 128
    A += 2 * torch.eye(nb, device="cuda")
    blk = pm.MatrixMult(A)
    xv = torch.randn(nb, device="cuda")
    yv = torch.empty(nb, device="cuda")
    ms = time_loop(lambda: blk.matvec(xv, out=yv), 50, 10)
    out["gemv_f32_4096_N"] = {"GB/s": gbs(4 * nb * nb, ms, 50), "us": ms / 50 * 1e3,
                              "note": "A (64 MiB) fits the 126 MB L2: >100% of HBM peak means L2 hits"}
    ms = time_loop(lambda: blk.rmatvec(xv, out=yv), 50, 10)
    out["gemv_f32_4096_T"] = {"GB/s": gbs(4 * nb * nb, ms, 50), "us": ms / 50 * 1e3}
    Op = pm.MPIBlockDiag([blk])
    xt = pm.DistributedArray(global_shape=nb * size, dtype=np.float32)
    xt.local_array.normal_()
    yd = Op.matvec(xt)
    x0 = xt.zeros_like()
    pm.cgls(Op, yd, x0=x0, niter=5, tol=0.0)
    torch.cuda.synchronize()
    comm.Barrier()
    t0 = time.perf_counter()
    xinv, istop, iit, r1, r2, cost = pm.cgls(Op, yd, x0=x0, niter=50, tol=0.0)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    err = (xinv - xt).norm()[0] / xt.norm()[0]
    out["cgls_blockdiag_4096_f32_50it"] = {"iters_per_s": 50 / dt, "ms_per_iter": dt / 50 * 1e3, "rel_err_vs_xtrue": float(err)}
    # HBM-bound GEMV (A = 1 GiB)
    A2 = torch.randn(32768, 8192, device="cuda")
    big = pm.MatrixMult(A2)
    xb, yb = torch.randn(8192, device="cuda"), torch.empty(32768, device="cuda")
    ms = time_loop(lambda: big.matvec(xb, out=yb), K, W)
    out["gemv_f32_32768x8192_N"] = {"GB/s": gbs(4 * A2.numel(), ms), "frac_hbm": gbs(4 * A2.numel(), ms) / hbm}
    xb2, yb2 = torch.randn(32768, device="cuda"), torch.empty(8192, device="cuda")
    ms = time_loop(lambda: big.rmatvec(xb2, out=yb2), K, W)
    out["gemv_f32_32768x8192_T"] = {"GB/s": gbs(4 * A2.numel(), ms), "frac_hbm": gbs(4 * A2.numel(), ms) / hbm}
    del A2, big
    # config 4 (i): bf16 matrix, single right-hand side (GEMV, HBM-bound)
    Ab = (torch.randn(32768, 16384, device="cuda") / 181).to(torch.bfloat16)
    bop = pm.MatrixMult(Ab)
    xb, yb = torch.randn(16384, device="cuda"), torch.empty(32768, device="cuda")
    ms = time_loop(lambda: bop.matvec(xb, out=yb), K, W)
    v = gbs(2 * Ab.numel(), ms)
    out["gemv_bf16_32768x16384_N"] = {"GB/s": v, "frac_hbm": v / hbm, "GF/s": 2 * Ab.numel() * K / (ms * 1e-3) / 1e9}
    # config 4 (ii): bf16 tile product on tensor cores, if the kernel is available
    try:
        m = n_ = k = 8192
        Am = (torch.randn(m, k, device="cuda") / 90).to(torch.bfloat16)
        Bm = (torch.randn(k, n_, device="cuda") / 90).to(torch.bfloat16)
        Cm = torch.empty(m, n_, device="cuda")

        def mm():
            L.check(L.lib.b2_gemm_bf16(L.ctx(), Am.data_ptr(), k, Bm.data_ptr(), n_, Cm.data_ptr(), n_, m, n_, k,
                                       L.OP_N, 0, L.stream()), "b2_gemm_bf16")
        ms = time_loop(mm, K, W)
        tf = 2.0 * m * n_ * k * K / (ms * 1e-3) / 1e12
        out["gemm_bf16_8192^3"] = {"TF/s": tf, "GF/s": tf * 1e3, "frac_tensor": tf / peaks.get("bf16_tflops", 1590.0)}
        del Am, Bm, Cm
    except Exception as exc:
        out["gemm_bf16_8192^3"] = {"unavailable": repr(exc)}
    del Ab, bop
    # config 4 through the operator: MPIMatrixMult SUMMA, 32768 x 32768 bf16 -> fp32, grid Pr x Pc
    try:
        grids = {1: (1, 1), 2: (1, 2), 4: (2, 2), 8: (2, 4)}
        if size in grids:
            Pr, Pc = grids[size]
            Ng = Kg = 32768
            ri, ci = divmod(rank, Pc)
            At = (torch.randn(Ng // Pr, Kg // Pc, device="cuda",
                              generator=torch.Generator(device="cuda").manual_seed(1 + rank)) / 181).to(torch.bfloat16)
            for Mg in (4096, 1):
                if Mg % Pc:
                    Mg = Pc
                for rep in (False, True):
                    Sop = pm.MPIMatrixMult(At, Mg, kind="summa", dtype="bfloat16", grid=(Pr, Pc), replicate=rep)
                    sizes = [(Kg // Pr) * (Mg // Pc)] * size
                    xs = pm.DistributedArray(global_shape=Kg * Mg, local_shapes=sizes, dtype=np.float32)
                    xs.local_array.normal_()
                    # sub-millisecond applies (M = 1 / Pc) need more launches to reach a steady state: with 5 timed
                    # steps the first, cold ones dominated (0.53 ms reported vs 0.35 ms steady, profiles/r01_diag_matmul_m1.json)
                    kk, ww = (5, 2) if Mg > 64 else (20, 10)
                    if Mg <= 64:        # let the power cap recover after the tensor-core runs (HBM-bound GEMV follows)
                        torch.cuda.synchronize()
                        time.sleep(0.5)
                    ms = time_loop(lambda: Sop.matvec(xs), kk, ww, comm)
                    fl = 2.0 * Ng * Kg * Mg
                    key = f"{'replicated' if rep else 'summa'}_bf16_32768_M{Mg}_grid{Pr}x{Pc}"
                    out[key] = {"TF/s": fl * kk / (ms * 1e-3) / 1e12, "ms": ms / kk,
                                "frac_tensor_total": fl * kk / (ms * 1e-3) / 1e12 / (size * peaks.get("bf16_tflops", 1590.0)),
                                "GB/s_A": 2.0 * Ng * Kg * kk / (ms * 1e-3) / 1e9,
                                "frac_hbm_A": 2.0 * Ng * Kg * kk / (ms * 1e-3) / 1e9 / (size * hbm)}
                    k2, w2 = (3, 1) if Mg > 64 else (10, 3)
                    ms = time_loop(lambda: Sop.rmatvec(Sop.matvec(xs)), k2, w2, comm)
                    out[key]["fwd+adj_ms"] = ms / k2
                    del Sop, xs
            del At
    except Exception as exc:
        out["summa_bf16_32768"] = {"error": repr(exc)}
    if size > 1:
        # BASELINE config 4 (i), the literal "32768-vec": 1-D row panels (grid P x 1, replicated mode) read every byte of A
        # exactly once per apply -- the 2-D grids above re-read A Pc times when M < Pc
        try:
            Ng = Kg = 32768
            At = (torch.randn(Ng // size, Kg, device="cuda",
                              generator=torch.Generator(device="cuda").manual_seed(1 + rank)) / 181).to(torch.bfloat16)
            Sop = pm.MPIMatrixMult(At, 1, kind="summa", dtype="bfloat16", grid=(size, 1), replicate=True)
            xs = pm.DistributedArray(global_shape=Kg, local_shapes=[Kg // size] * size, dtype=np.float32)
            xs.local_array.normal_()
            torch.cuda.synchronize()
            time.sleep(0.5)
            ms = time_loop(lambda: Sop.matvec(xs), 20, 10, comm)
            gb = 2.0 * Ng * Kg * 20 / (ms * 1e-3) / 1e9
            out[f"replicated_bf16_32768_M1_grid{size}x1"] = {"ms": ms / 20, "GB/s_A": gb, "frac_hbm_A": gb / (size * hbm),
                                                             "GF/s": 2.0 * Ng * Kg * 20 / (ms * 1e-3) / 1e9}
            ms = time_loop(lambda: Sop.rmatvec(Sop.matvec(xs)), 10, 3, comm)
            out[f"replicated_bf16_32768_M1_grid{size}x1"]["fwd+adj_ms"] = ms / 10
            del Sop, xs, At
        except Exception as exc:
            out[f"replicated_bf16_32768_M1_grid{size}x1"] = {"error": repr(exc)}
    # --- config 5: Fredholm1 (64 slices per GPU, 256 x 256 x 64, complex64) -------------------
    nsl, ns, nr, nv = 64, 256, 256, 64
    G = torch.randn(nsl, ns, nr, device="cuda", dtype=torch.complex64)
    Fr = pm.MPIFredholm1(G, nz=nv, dtype=np.complex64, fused=False)
    xm = pm.DistributedArray(global_shape=nsl * size * nr * nv, partition=pm.Partition.BROADCAST, dtype=np.complex64)
    xm.local_array.normal_()
    ms = time_loop(lambda: Fr.matvec(xm), K, W, comm)
    fl = 8.0 * nsl * size * ns * nr * nv
    out["fredholm1_c64_64x256x256x64_per_gpu"] = {"GF/s": fl * K / (ms * 1e-3) / 1e9, "us": ms / K * 1e3,
                                                  "mode": "product kernel + NCCL all-gather"}
    if size > 1:
        Ff = pm.MPIFredholm1(G, nz=nv, dtype=np.complex64, fused=True)
        ms = time_loop(lambda: Ff.matvec(xm), K, W, comm)
        out["fredholm1_fused_peer_c64_64x256x256x64_per_gpu"] = {
            "GF/s": fl * K / (ms * 1e-3) / 1e9, "us": ms / K * 1e3,
            "mode": "ONE kernel: product + all-gather via P2P stores into IPC-mapped peer buffers"}
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()

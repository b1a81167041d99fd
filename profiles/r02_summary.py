"""Round-2 evidence reducer (run here, no GPU): turns what the GPU runs left in gpurun_out/ into tracked summaries.

    python profiles/r02_summary.py

writes profiles/r02_ncu_summary.md (per-launch durations of the Fredholm path, the `ncu --set full` metrics of
fredholm_tc_kernel) and copies the bench lines to profiles/r02_bench_n*.json.
"""
import csv
import collections
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
sys.path.insert(0, os.path.join(ROOT, "profiles"))
from summarise import read_report, to_bytes, to_us  # noqa: E402


def launches(path):
    rows = [r for r in csv.reader(ln for ln in open(path) if not ln.startswith("=="))]
    hdr = rows[0]
    ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
    d = collections.OrderedDict()
    for r in rows[1:]:
        d.setdefault(r[ki], []).append(float(r[vi].replace(",", "")) / 1e3)
    return d


def main():
    md = ["# Round 2 -- ncu evidence\n"]
    p = os.path.join(OUT, "fr_launches_pack1.csv")        # launch list taken after the single-pass pack kernel landed
    if not os.path.exists(p):
        p = os.path.join(OUT, "fr_launches_h2.csv")
    if os.path.exists(p):
        md.append("## Launch list of the MPIFredholm1 tensor-core path (config 5: 64 slices of 256 x 256 x 64 complex64)\n")
        md.append("`ncu --metrics gpu__time_duration.sum --clock-control none` over `profiles/fredholm_tc_check.py --time "
                  "--time-only` (per-launch times under ncu are cold-cache and serialised: shares, not absolutes).\n")
        md.append("| kernel | launches | mean us | min us |\n|---|---|---|---|")
        for k, v in launches(p).items():
            md.append(f"| `{k[:90]}` | {len(v)} | {sum(v) / len(v):.2f} | {min(v):.2f} |")
        shutil.copy(p, os.path.join(ROOT, "profiles", "r02_launches_fredholm.csv"))
        md.append("")
    for tag in ("h2", "b3"):
        rep = os.path.join(OUT, f"r02_fredholm_tc_{tag}.ncu-rep")
        if not os.path.exists(rep):
            continue
        md.append(f"## `ncu --set full` of fredholm_tc_kernel ({'fp16x2' if tag == 'h2' else 'bf16x3'} mode)\n")
        md.append("| launch | duration us | DRAM read MB | DRAM write MB | DRAM % of ncu peak | tensor pipe % | warps active % | "
                  "L2 MB | regs | grid x block | dyn smem |\n|---|---|---|---|---|---|---|---|---|---|---|")
        for i, d in enumerate(read_report(rep)):
            def g(k, f=lambda v, u: v):
                return f(*d[k]) if k in d else float("nan")
            md.append(f"| {i} `{d['kernel'][:40]}` | {g('duration', to_us):.2f} | {g('dram_read', to_bytes) / 1e6:.1f} | "
                      f"{g('dram_write', to_bytes) / 1e6:.1f} | {g('dram_pct_of_ncu_peak')} | {g('tensor_pipe_pct')} | "
                      f"{g('warps_active_pct')} | {g('l2_bytes', to_bytes) / 1e6:.1f} | {g('regs')} | {g('grid')} x {g('block')} | "
                      f"{g('dyn_smem')} |")
        md.append("")
    for n in (1, 2, 4, 8):
        src = os.path.join(OUT, f"r02_bench_n{n}.json")
        if os.path.exists(src) and os.path.getsize(src) > 10:
            try:        # a line whose CGLS section fell back to eager iterations is a diagnostic, not evidence
                line = json.loads(open(src).read().strip().splitlines()[-1])
                bad = (line.get("extra", {}).get("cgls_blockdiag_4096_f32_50it", {}) or {}).get("cuda_graph_error")
            except Exception:
                bad = "unreadable"
            if bad:
                print(f"NOT copying {src}: {bad}")
                continue
            shutil.copy(src, os.path.join(ROOT, "profiles", f"r02_bench_n{n}.json"))
    for name in ("r02_multi2.log", "r02_multi8.log", "r02_pytest_gpu.log"):
        src = os.path.join(OUT, name)
        if os.path.exists(src):
            with open(src) as f:
                tail = f.read()[-3000:]
            with open(os.path.join(ROOT, "profiles", name), "w") as f:
                f.write(tail)
    with open(os.path.join(ROOT, "profiles", "r02_ncu_summary.md"), "w") as f:
        f.write("\n".join(md) + "\n")
    print("\n".join(md)[:4000])


if __name__ == "__main__":
    main()

"""Turn the ncu reports brought back in gpurun_out/ into the tracked summaries
under profiles/ (run here, no GPU needed):

    python profiles/summarise.py r01

writes profiles/<round>_ncu_summary.md, profiles/<round>_launches_bench.csv and
profiles/traffic.json (DRAM bytes per launch of the headline kernel, read by bench.py).
"""
import csv
import io
import json
import os
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "gpurun_out")
WANT = OrderedDict([
    ("gpu__time_duration.sum", "duration"),
    ("dram__bytes_read.sum", "dram_read"),
    ("dram__bytes_write.sum", "dram_write"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram_pct_of_ncu_peak"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor_pipe_pct"),
    ("sm__inst_executed_pipe_tensor.sum", "tensor_inst"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps_active_pct"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm_pct"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__shared_mem_per_block_dynamic", "dyn_smem"),
])


def to_bytes(val, unit):
    v = float(val.replace(",", ""))
    mult = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}.get(unit, 1)
    return v * mult


def to_us(val, unit):
    v = float(val.replace(",", ""))
    return v * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(unit, 1.0)


def read_report(path):
    raw = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(raw)))
    hdr, units = rows[0], rows[1]
    out = []
    for r in rows[2:]:
        d = {"kernel": r[hdr.index("Kernel Name")]}
        for k, short in WANT.items():
            if k in hdr:
                i = hdr.index(k)
                d[short] = (r[i], units[i])
        out.append(d)
    return out


def main():
    rnd = sys.argv[1] if len(sys.argv) > 1 else "r01"
    lines = [f"# ncu summaries, round {rnd}", "",
             "Source: `ncu --set full --clock-control none --import-source on` captures taken on the B200 box by",
             "`profiles/capture.sh`; the binary reports stay in gpurun_out/ (scratch).  Durations under ncu are",
             "cold-cache / serialised: use them for shares and traffic, not as benchmark numbers.", ""]
    traffic = {}
    for rep in ("prof_stencil", "prof_extras", "prof_gemm", "prof_fused"):
        path = os.path.join(OUT, rep + ".ncu-rep")
        if not os.path.exists(path):
            continue
        lines += [f"## {rep}.ncu-rep", "",
                  "| kernel | us | DRAM read | DRAM write | DRAM % (ncu peak) | tensor pipe % | warps active % | regs | grid x block |",
                  "|---|---|---|---|---|---|---|---|---|"]
        for d in read_report(path):
            name = d["kernel"].replace("void <unnamed>::", "").split("(")[0]
            us = to_us(*d["duration"])
            rd, wr = to_bytes(*d["dram_read"]), to_bytes(*d["dram_write"])
            lines.append(f"| `{name}` | {us:.1f} | {rd / 1e6:.1f} MB | {wr / 1e6:.1f} MB | {float(d['dram_pct_of_ncu_peak'][0]):.1f} | "
                         f"{float(d.get('tensor_pipe_pct', ('0', ''))[0]):.1f} | {float(d['warps_active_pct'][0]):.1f} | "
                         f"{d['regs'][0]} | {d['grid'][0]} x {d['block'][0]} |")
            if "stencil_vec_kernel<float, 10" in d["kernel"]:
                traffic.setdefault("stencil_samples", []).append(rd + wr)
        lines.append("")
    if traffic.get("stencil_samples"):
        s = traffic.pop("stencil_samples")
        traffic["stencil_vec_kernel_f32_centered3_bytes_per_launch"] = sum(s) / len(s)
        traffic["source"] = f"{rnd}: dram__bytes_read.sum + dram__bytes_write.sum, mean of {len(s)} launches (32768 x 8192 f32)"
        with open(os.path.join(ROOT, "profiles", "traffic.json"), "w") as f:
            json.dump(traffic, f, indent=1)
    # launch list of the bench command: per-kernel totals and shares
    lpath = os.path.join(OUT, "launches_bench.csv")
    if os.path.exists(lpath):
        txt = [ln for ln in open(lpath) if ln.startswith('"')]
        rows = list(csv.reader(io.StringIO("".join(txt))))
        hdr = rows[0]
        ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
        agg = OrderedDict()
        for r in rows[1:]:
            if len(r) <= vi:
                continue
            name = r[ki].replace("void <unnamed>::", "").split("(")[0][:90]
            us = to_us(r[vi], r[ui])
            a = agg.setdefault(name, [0, 0.0])
            a[0] += 1
            a[1] += us
        tot = sum(a[1] for a in agg.values())
        with open(os.path.join(ROOT, "profiles", f"{rnd}_launches_bench.csv"), "w") as f:
            f.write("kernel,launches,total_us,share_pct\n")
            for name, (cnt, us) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
                f.write(f"\"{name}\",{cnt},{us:.1f},{100 * us / tot:.2f}\n")
        lines += ["## launch list of `python bench.py --steps 3 --warmup 3` (first 600 launches)", "",
                  f"see `{rnd}_launches_bench.csv` (per-kernel launch counts, summed gpu__time_duration, share)", ""]
    with open(os.path.join(ROOT, "profiles", f"{rnd}_ncu_summary.md"), "w") as f:
        f.write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    main()

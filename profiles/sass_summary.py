"""`cuobjdump -sass` of libb200lops.so reduced to per-kernel counts of the instructions that prove the tensor-core /
TMA / TMEM / peer-memory paths.  Runs on the build host (no GPU):

    python profiles/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "pylops_mpi_b200", "libb200lops.so")
COLS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTCBAR", "LDTM", "SYNCS", "ATOMG", "MEMBAR.SC.SYS", "CCTL.IVALL"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                           text=True).stdout.splitlines()
    archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
    counts, order, cur, k = {}, [], None, -1
    for ln in sass.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            k += 1
            cur = names[k] if k < len(names) else m.group(1)
            cur = re.sub(r"\(anonymous namespace\)::", "", cur)
            cur = re.sub(r"\(.*$", "", cur).strip()
            if cur not in counts:
                counts[cur] = collections.Counter()
                order.append(cur)
            continue
        if cur is None:
            continue
        m = re.search(r"^\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
        if not m:
            continue
        op = m.group(1)
        c = counts[cur]
        if op.startswith("UTCHMMA"):
            c["UTCHMMA.2CTA" if ".2CTA" in op else "UTCHMMA"] += 1
        for key in ("UTMALDG", "UTCBAR", "LDTM", "SYNCS", "ATOMG"):
            if op.startswith(key):
                c[key] += 1
        if op.startswith("MEMBAR.SC.SYS"):
            c["MEMBAR.SC.SYS"] += 1
        if op.startswith("CCTL.IVALL"):
            c["CCTL.IVALL"] += 1
    print(f"# cuobjdump -sass pylops_mpi_b200/libb200lops.so: {len(order)} kernels; architectures in the fatbin: {archs}")
    print("# per-kernel instruction counts (kernels that use the tensor / TMA / TMEM / peer-memory instructions)\n")
    print("kernel | " + " | ".join(COLS))
    for name in order:
        c = counts[name]
        if any(c[col] for col in COLS if col not in ("SYNCS",)):
            print(name + " | " + " | ".join(str(c[col]) for col in COLS))
    plain = [n for n in order if not any(counts[n][col] for col in COLS if col != "SYNCS")]
    print(f"\n# {len(plain)} further kernels use none of these (SIMT / HBM-bound paths), e.g.: " + ", ".join(sorted(set(
        re.sub(r"<.*", "", n).replace("void ", "") for n in plain))[:40]))


if __name__ == "__main__":
    sys.exit(main())

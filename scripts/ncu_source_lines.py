"""Attribute the warp-stall samples of an ncu report to CUDA source lines.

    python scripts/ncu_source_lines.py gpurun_out/prof_workers_r1.ncu-rep k_search_workersILi2E [min_pct]

ncu's `--page source --csv` lists SASS only; the line table comes from `nvdisasm -g` on the cubin of the
in-tree libpomcp_b200.so (must be the build that was profiled).  Prints per-line sample share, warp
instructions executed and the two dominant stall reasons, then totals per file region."""
import collections
import csv
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "pomcp.jl_b200", "libpomcp_b200.so")
CSRC = os.path.join(ROOT, "pomcp.jl_b200", "csrc")


def sass_lines(mangled_part):
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", SO], cwd=td, check=True, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        txt = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    out, cur, on = {}, None, False
    for ln in txt.split("\n"):
        if ln.startswith("//----") and ".text." in ln:
            on = mangled_part in ln
            continue
        if not on:
            continue
        m = re.match(r'\s*//## File "([^"]+)", line (\d+)', ln)
        if m:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", ln)
        if m:
            out[int(m.group(1), 16)] = (m.group(2).strip(), cur)
    return out


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    min_pct = float(sys.argv[3]) if len(sys.argv) > 3 else 0.3
    sass = sass_lines(kern)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(data[0][0], 16)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    agg, inst, st = collections.Counter(), collections.Counter(), collections.defaultdict(collections.Counter)
    tinst = collections.Counter()
    for r in data:
        off = int(r[0], 16) - base
        loc = sass.get(off, (None, None))[1]
        n = int(r[ix["# Samples"]])
        agg[loc] += n
        inst[loc] += int(r[ix["Instructions Executed"]])
        tinst[loc] += int(r[ix["Thread Instructions Executed"]])
        for s in stalls:
            st[loc][s] += int(r[ix[s]])
    tot = sum(agg.values())
    src = {f: open(os.path.join(CSRC, f)).read().split("\n") for f in os.listdir(CSRC)}
    print(f"total samples {tot}, warp instructions {sum(inst.values())}")
    tot_inst = sum(inst.values())
    by_inst = "--by-inst" in sys.argv  # rank the lines by warp instructions executed instead of by stall samples
    for loc, n in sorted(agg.items(), key=lambda x: (x[0] or ("", 0))):
        if (inst[loc] < tot_inst * min_pct / 100) if by_inst else (n < tot * min_pct / 100):
            continue
        f, l = loc if loc else ("?", 0)
        text = src[f][l - 1].strip()[:88] if f in src and 0 < l <= len(src[f]) else ""
        top = ", ".join(f"{k[6:]}:{v * 100 // max(n, 1)}%" for k, v in st[loc].most_common(2))
        lanes = tinst[loc] / max(inst[loc], 1)
        print(f"{f}:{l}".ljust(20), f"{n * 100 / tot:5.1f}%", str(inst[loc]).rjust(12), f"{lanes:5.1f}", top.ljust(36), text)
    byfile = collections.Counter()
    for loc, n in agg.items():
        byfile[loc[0] if loc else "?"] += n
    print({k: round(v * 100 / tot, 1) for k, v in byfile.most_common()})


if __name__ == "__main__":
    main()

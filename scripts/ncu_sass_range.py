"""Per-SASS-instruction stall samples of a source-line range (the batched descent loop, say).

    python scripts/ncu_sass_range.py gpurun_out/prof_workers_r1.ncu-rep k_search_workersILi2ELi2ELi2ELb0 1051 1230

Prints, in address order, every instruction between the first and the last one attributed to
tree.cuh:[lo, hi]: samples, executions, samples per 1000 executions (proportional to the cycles a
warp spends at that instruction) and the dominant stall reason."""
import csv
import subprocess
import sys

from ncu_source_lines import sass_lines


def main():
    rep, kern, lo, hi = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4])
    sass = sass_lines(kern)
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(txt.splitlines()))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[2:] if len(r) == len(hdr)]
    base = int(data[0][0], 16)
    stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    offs = [o for o, (_, loc) in sass.items() if loc and loc[0] == "tree.cuh" and lo <= loc[1] <= hi]
    a0, a1 = min(offs), max(offs)
    tot = 0
    for r in data:
        off = int(r[0], 16) - base
        if not (a0 <= off <= a1):
            continue
        n, ex = int(r[ix["# Samples"]]), int(r[ix["Instructions Executed"]])
        tot += n
        top = max(stalls, key=lambda s: int(r[ix[s]]))
        ins, loc = sass.get(off, ("?", None))
        print(f"{off:05x} {n:7d} {ex:10d} {n * 1000 / max(ex, 1):7.1f} {top[6:18]:12s} {ins[:64]:64s} "
              f"{(loc[0][:8] + ':' + str(loc[1])) if loc else ''}")
    print("samples in range", tot)


if __name__ == "__main__":
    main()

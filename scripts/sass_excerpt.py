"""SASS of a source-line range of a kernel of the in-tree library, with the line each instruction comes from
(nvdisasm -g line table).  Writes the excerpts DESIGN.md cites:

    python scripts/sass_excerpt.py k_search_workersILi2ELi2ELi2ELb0 tree.cuh 305 335 profiles/r2_sass_rollout_block.txt
    python scripts/sass_excerpt.py k_search_workersILi2ELi2ELi2ELb0 tree.cuh 1262 1420 profiles/r2_sass_descent_level.txt
"""
import collections
import sys

from ncu_source_lines import sass_lines


def main():
    kern, fname, lo, hi, out = sys.argv[1], sys.argv[2], int(sys.argv[3]), int(sys.argv[4]), sys.argv[5]
    GAP = int(sys.argv[6], 0) if len(sys.argv) > 6 else 0x200  # largest address gap inside the excerpt
    sass = sass_lines(kern)
    offs = sorted(o for o, (_, loc) in sass.items() if loc and loc[0] == fname and lo <= loc[1] <= hi)
    # the longest run of addresses whose gaps are small (the loop body; stray inlined copies elsewhere are dropped)
    runs, cur = [], [offs[0]]
    for o in offs[1:]:
        if o - cur[-1] <= GAP:
            cur.append(o)
        else:
            runs.append(cur)
            cur = [o]
    runs.append(cur)
    best = max(runs, key=len)
    a0, a1 = best[0], best[-1]
    ops = collections.Counter()
    with open(out, "w") as f:
        f.write(f"# cuobjdump/nvdisasm -g of pomcp.jl_b200/libpomcp_b200.so, kernel *{kern}*, the contiguous address range that\n"
                f"# holds {fname}:{lo}-{hi} (instructions inlined from other lines / files in between are kept)\n")
        n = 0
        for o in sorted(sass):
            if a0 <= o <= a1:
                ins, loc = sass[o]
                ops[ins.split()[0].split(".")[0] if not ins.startswith("@") else ins.split()[1].split(".")[0]] += 1
                f.write(f"/*{o:05x}*/ {ins:72s} // {loc[0] + ':' + str(loc[1]) if loc else ''}\n")
                n += 1
        f.write(f"# {n} instructions; by opcode: " + ", ".join(f"{k} {v}" for k, v in ops.most_common()) + "\n")
    print(out, n, "instructions", dict(ops.most_common(12)))


if __name__ == "__main__":
    main()

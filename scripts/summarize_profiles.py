"""Turn the ncu artefacts brought back in gpurun_out/ into the tracked summaries under profiles/.

    python scripts/summarize_profiles.py r1 [steps_per_decision]

Reads gpurun_out/launches_<tag>.csv (the `--metrics gpu__time_duration.sum` launch list of
bench.py) and gpurun_out/prof_*_<tag>.ncu-rep (`--set full` captures), writes
profiles/<tag>_launches_summary.csv, profiles/<tag>_ncu_<name>.csv and updates
profiles/roofline_constants.json (the constants bench.py folds into its roofline objects)."""
import collections
import csv
import glob
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GP = os.path.join(ROOT, "gpurun_out")

KEEP = [
    "Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
    "smsp__thread_inst_executed_per_inst_executed.ratio", "smsp__thread_inst_executed.sum", "dram__bytes_read.sum",
    "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sectors_op_atom.sum", "lts__t_sectors_op_red.sum",
    "smsp__average_warp_latency_per_inst_issued.ratio",
    "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_membar_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
    "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
]


def to_ns(v, unit):
    return float(v.replace(",", "")) * {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(unit, 1)


def launches(tag):
    path = os.path.join(GP, f"launches_{tag}.csv")
    if not os.path.exists(path):
        return None
    lines = [ln for ln in open(path) if not ln.startswith("==")]
    tot, cnt = collections.Counter(), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "")
        tot[name] += to_ns(row["Metric Value"], row["Metric Unit"])
        cnt[name] += 1
    total = sum(tot.values())
    with open(os.path.join(OUT, f"{tag}_launches_summary.csv"), "w") as f:
        f.write("# ncu --metrics gpu__time_duration.sum --clock-control none python bench.py --steps 3 --warmup 3 "
                "--no-cpu-baseline (per-launch times are cold-cache and serialised: compare SHARES)\n")
        f.write("kernel,launches,total_ms,share_pct,avg_us\n")
        for k, v in tot.most_common():
            f.write(f"{k},{cnt[k]},{v / 1e6:.3f},{v / total * 100:.2f},{v / cnt[k] / 1e3:.2f}\n")
    return {k: (cnt[k], v) for k, v in tot.items()}


def ncu_raw(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    return rows[0], rows[1], rows[2:]


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r1"
    steps_per_decision = float(sys.argv[2]) if len(sys.argv) > 2 else None
    if steps_per_decision is None:  # rollout steps of the profiled decision (-s 1: the second), printed by probe.py step
        log = os.path.join(GP, f"{tag}_ncu_search.log")
        if os.path.exists(log):
            dec = [json.loads(ln) for ln in open(log) if ln.startswith("{")]
            if len(dec) >= 2:
                steps_per_decision = float(dec[1]["rollout_steps"])
    os.makedirs(OUT, exist_ok=True)
    consts_path = os.path.join(OUT, "roofline_constants.json")
    consts = json.load(open(consts_path)) if os.path.exists(consts_path) else {}
    launches(tag)
    for rep in sorted(glob.glob(os.path.join(GP, f"prof_*_{tag}.ncu-rep"))):
        name = os.path.basename(rep)[:-len(".ncu-rep")]
        hdr, units, rows = ncu_raw(rep)
        cols = [c for c in KEEP if c in hdr]
        with open(os.path.join(OUT, f"{tag}_ncu_{name}.csv"), "w") as f:
            f.write(f"# ncu --set full --clock-control none --import-source on ... -> {name}.ncu-rep (gpurun_out/, not tracked)\n")
            w = csv.writer(f)
            w.writerow(["metric", "unit"] + [f"launch{i}" for i in range(len(rows))])
            for c in cols:
                i = hdr.index(c)
                w.writerow([c, units[i]] + [r[i] for r in rows])
        for r in rows:
            kn = r[hdr.index("Kernel Name")]
            g = lambda m: float(r[hdr.index(m)].replace(",", ""))  # noqa: E731
            if "k_search_workers" in kn:
                ti = g("smsp__inst_executed.sum") * g("smsp__thread_inst_executed_per_inst_executed.ratio")
                consts["workers_issue_active_pct"] = g("smsp__issue_active.avg.pct_of_peak_sustained_active")
                consts["rollout_warp_efficiency"] = g("smsp__thread_inst_executed_per_inst_executed.ratio") / 32.0
                consts["workers_thread_inst_per_launch"] = ti
                sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                consts["workers_dram_bytes_per_launch"] = (
                    g("dram__bytes_read.sum") * sc.get(units[hdr.index("dram__bytes_read.sum")], 1) +
                    g("dram__bytes_write.sum") * sc.get(units[hdr.index("dram__bytes_write.sum")], 1))
                if "lts__t_sectors_op_atom.sum" in hdr and "lts__t_sectors_op_red.sum" in hdr:
                    consts["workers_l2_atomic_sectors_per_launch"] = g("lts__t_sectors_op_atom.sum") + g("lts__t_sectors_op_red.sum")
                if steps_per_decision:
                    consts["rollout_thread_inst_per_step"] = ti / steps_per_decision
                    consts["rollout_steps_per_profiled_launch"] = steps_per_decision
            if "k_pf_" in kn:
                unit_r = units[hdr.index("dram__bytes_read.sum")]
                scale = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
                consts["pf_dram_bytes_per_update"] = (g("dram__bytes_read.sum") * scale.get(unit_r, 1) +
                                                      g("dram__bytes_write.sum") *
                                                      scale.get(units[hdr.index("dram__bytes_write.sum")], 1))
    # Single-pass instruction counts + the step counters of the SAME launches (scripts/gpu_session.sh ncu: three SM
    # counters fit one pass; `probe.py step --repeat 8` prints one JSON line per decision).  The instruction count of a
    # launch depends on the tree it grows (how long its rollouts live), so it is modelled as
    #     thread instructions = a * simulations + b * live rollout steps
    # fitted by least squares over all the launches of the run, residuals kept; bench.py applies (a, b) to the
    # simulations and steps it counts itself.
    one = os.path.join(GP, f"workers_inst_{tag}.csv")
    log1 = os.path.join(GP, f"workers_inst_{tag}.log")
    if os.path.exists(one) and os.path.exists(log1):
        import numpy as np
        dec = [json.loads(ln) for ln in open(log1) if ln.startswith("{")]
        per = collections.defaultdict(dict)
        for row in csv.DictReader([ln for ln in open(one) if not ln.startswith("==")]):
            per[int(row["ID"])][row["Metric Name"]] = float(row["Metric Value"].replace(",", ""))
        ids = sorted(per)
        if len(ids) == len(dec) and len(ids) >= 2:
            ti = np.array([per[i]["smsp__thread_inst_executed.sum"] for i in ids])
            wi = np.array([per[i]["smsp__inst_executed.sum"] for i in ids])
            sims = np.array([d["simulations"] for d in dec], dtype=float)
            steps = np.array([d["rollout_steps"] for d in dec], dtype=float)
            A = np.stack([sims, steps], axis=1)
            (ca, cb), *_ = np.linalg.lstsq(A, ti, rcond=None)
            rel = (A @ np.array([ca, cb]) - ti) / ti
            consts["workers_thread_inst_per_sim"] = float(ca)
            consts["workers_thread_inst_per_rollout_step"] = float(cb)
            consts["workers_inst_fit"] = {"launches": len(ids), "rms_rel_residual": float(np.sqrt((rel ** 2).mean())),
                                          "max_rel_residual": float(np.abs(rel).max()),
                                          "from": f"gpurun_out/workers_inst_{tag}.csv + .log (one ncu pass per launch)"}
            consts["workers_inst_samples"] = [[float(t), float(st), float(sm)] for t, st, sm in zip(ti, steps, sims)]
            consts["rollout_warp_efficiency"] = float((ti / wi / 32.0).mean())
            consts["workers_issue_active_pct"] = float(np.mean([per[i]["smsp__issue_active.avg.pct_of_peak_sustained_active"]
                                                                for i in ids]))
            consts["workers_thread_inst_per_launch"] = float(ti.mean())
            with open(os.path.join(OUT, f"{tag}_workers_inst_per_launch.csv"), "w") as f:
                f.write("# ncu --metrics smsp__thread_inst_executed.sum,smsp__inst_executed.sum,smsp__issue_active... "
                        "--clock-control none -k regex:k_search_workers -c 8 python scripts/probe.py step --repeat 8\n")
                f.write("launch,thread_inst,warp_inst,issue_active_pct,simulations,rollout_steps,tree_steps,model_thread_inst,rel_residual\n")
                for k, i in enumerate(ids):
                    f.write(f"{i},{ti[k]:.0f},{wi[k]:.0f},{per[i]['smsp__issue_active.avg.pct_of_peak_sustained_active']},"
                            f"{sims[k]:.0f},{steps[k]:.0f},{dec[k]['tree_steps']},{ca * sims[k] + cb * steps[k]:.0f},{rel[k]:.5f}\n")
                f.write(f"# fit: thread_inst = {ca:.1f} * simulations + {cb:.3f} * rollout_steps; rms relative residual "
                        f"{np.sqrt((rel ** 2).mean()):.4f}\n")
    # pipe utilisation of the worker kernel (which pipe binds)
    pp = os.path.join(GP, f"workers_pipes_{tag}.csv")
    if os.path.exists(pp):
        pipes = {}
        for row in csv.DictReader([ln for ln in open(pp) if not ln.startswith("==")]):
            m = row["Metric Name"]
            if m.startswith("sm__inst_executed_pipe_"):
                pipes[m[len("sm__inst_executed_pipe_"):].split(".")[0]] = float(row["Metric Value"])
            elif m.startswith("sm__warps_active"):
                pipes["warps_active"] = float(row["Metric Value"])
            elif m.startswith("smsp__issue_active"):
                pipes["issue_active"] = float(row["Metric Value"])
        consts["workers_pipe_pct"] = pipes
        with open(os.path.join(OUT, f"{tag}_workers_pipes.csv"), "w") as f:
            f.write("# ncu --metrics sm__inst_executed_pipe_*.avg.pct_of_peak_sustained_active ... -k regex:k_search_workers -s 1 -c 1 "
                    "python scripts/probe.py step\nmetric,pct_of_peak_sustained_active\n")
            for k, v in sorted(pipes.items(), key=lambda kv: -kv[1]):
                f.write(f"{k},{v}\n")
    consts["source"] = f"scripts/summarize_profiles.py {tag}"
    json.dump(consts, open(consts_path, "w"), indent=1)
    print(json.dumps(consts, indent=1))


if __name__ == "__main__":
    main()

"""One parametrised measurement tool for the GPU box (every table of DESIGN.md comes out of a subcommand):

    python scripts/probe.py oracle-values --queries 1000000 --seeds 12 [--rollouts 64 --procs 6] --out profiles/r2_oracle_values_1e6.json
        sequential oracle (CPU, deterministic: can run anywhere), RockSample(7,8), initial belief, per seed
    python scripts/probe.py knob --queries 1000000 --seeds 12 --knob inflight_max --values 64,256,1184
        batched search against the oracle per value of one solver knob (inflight_max / inflight_min /
        inflight_ramp_div / rollouts_per_leaf / n_trees): action agreement, best-arm V, decision time
    python scripts/probe.py equal-time --budget-ms 14 --seeds 12 --values 64,256,1184
        the same at equal WALL TIME per decision: the query budget of every in-flight cap is what fits the time
    python scripts/probe.py cycles [--queries Q] [--inflight 1184] [--rollouts 64]
        where a batched worker's cycles go (pomcp_worker_cycles)
    python scripts/probe.py sir --problem rs78|lightdark --log2 12,20,24,26
        SIR update time and algorithmic GB/s per particle count and path
    python scripts/probe.py rates
        batched search rates of the concrete problems with the library defaults
    python scripts/probe.py c4      (or: python -m torch.distributed.run --nproc-per-node 2 ... scripts/probe.py c4)
        BASELINE configs[3] (LightDark1D, 1e6 particles) on 1 / 2 GPUs: what the second GPU buys
    python scripts/probe.py step [--queries Q] [--rollouts R] [--trees T]
        one warmed-up decision (search + SIR 2^20) on the bench workload: the ncu target
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def rs78():
    import pomcp_jl_b200 as pj
    return pj.RockSamplePOMDP(7, 8, seed=7008)


def _oracle_one(job):
    Q, seed, c, R = job
    from oracle import oracle as O
    from pomcp_jl_b200 import _abi as abi
    op = O.OraclePlanner(rs78(), O.default_params(c=c, seed=seed, tree_queries=Q, belief_mode=abi.BELIEF_UNWEIGHTED,
                                                  rollouts_per_leaf=R))
    op.set_belief_initial()
    t = time.time()
    rc, a, N, V = op.search(Q)
    assert rc == 0
    out = (a, V, N, time.time() - t, op.counters())
    op.close()
    return out


def oracle_values(Q, seeds, cache=None, c=20.0, R=1, procs=1):
    """[(action, V[nA])] of the sequential oracle per seed; read from `cache` (a file this tool wrote) when it fits."""
    if cache and cache != "none" and os.path.exists(cache):
        d = json.load(open(cache))
        if d["queries"] == Q and d["c"] == c and d.get("rollouts_per_leaf", 1) == R and len(d["runs"]) >= len(seeds):
            return [(r["action"], np.array(r["V"])) for r in d["runs"][:len(seeds)]]
    jobs = [(Q, seed, c, R) for seed in seeds]
    if procs > 1:
        import multiprocessing as mp
        with mp.get_context("spawn").Pool(procs) as pool:
            return pool.map(_oracle_one, jobs, chunksize=1)
    return [_oracle_one(j) for j in jobs]


def cmd_oracle_values(a):
    R = int(a.rollouts.split(",")[0]) or 1
    runs = oracle_values(a.queries, range(a.seeds), R=R, procs=a.procs)
    d = {"problem": "RockSample(7,8) rocks_seed 7008, initial belief, unweighted", "c": 20.0, "queries": a.queries,
         "rollouts_per_leaf": R,
         "runs": [{"seed": s, "action": int(r[0]), "V": [float(x) for x in r[1]], "N": [int(x) for x in r[2]],
                   "seconds": r[3], "rollout_steps": r[4]["rollout_steps"]} for s, r in enumerate(runs)]}
    vo = np.array([r[1][r[0]] for r in runs])
    print(f"Q={a.queries}: oracle best-arm V {vo.mean():.3f} +- {vo.std(ddof=1) / np.sqrt(len(vo)):.3f} (SE), actions "
          f"{[int(r[0]) for r in runs]}")
    if a.out:
        json.dump(d, open(a.out, "w"))


def gpu_runs(Q, seeds, **kw):
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    pomdp = rs78()
    res = []
    for seed in seeds:
        solver = pj.POMCPSolver(c=20.0, rng=seed, tree_queries=Q, search_mode="batched",
                                max_nodes=4 * Q * kw.get("n_trees", 1) + 65536, **kw)
        gp = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
        gp.set_belief_initial()
        a, N, V = gp.search(Q)
        ms = gp.phase_ms(abi.PHASE_SEARCH)[0]
        cs = gp.counters()
        gp.close()
        res.append((a, V, ms, cs))
    return res


def report(tag, runs, orc):
    ag = [r[0] == o[0] for r, o in zip(runs, orc)]
    vg = np.array([r[1][r[0]] for r in runs])
    vo = np.array([o[1][o[0]] for o in orc])
    # value the batched tree assigns to the ORACLE's action, against the oracle's own value of it
    dv = np.array([r[1][o[0]] - o[1][o[0]] for r, o in zip(runs, orc)])
    ms = np.mean([r[2] for r in runs])
    sims = np.mean([r[3]["simulations"] for r in runs])  # all trees of the decision
    steps = np.mean([r[3]["rollout_steps"] for r in runs])
    line = {"setting": tag, "action_eq_oracle": f"{sum(ag)}/{len(ag)}", "best_V": float(vg.mean()),
            "best_V_se": float(vg.std(ddof=1) / np.sqrt(len(vg))), "oracle_best_V": float(vo.mean()),
            "oracle_best_V_se": float(vo.std(ddof=1) / np.sqrt(len(vo))), "dV_on_oracle_action": float(dv.mean()),
            "dV_se": float(dv.std(ddof=1) / np.sqrt(len(dv))), "decision_ms": float(ms), "sims": float(sims),
            "sims_per_s": float(sims / ms * 1e3), "rollout_steps_per_s": float(steps / ms * 1e3)}
    print(json.dumps(line), flush=True)
    return line


def cmd_knob(a):
    seeds = range(a.seeds)
    orc = oracle_values(a.queries, seeds, a.oracle_cache)
    rows = []
    for v in a.values.split(","):
        if a.knob == "n_trees":  # the same total budget split over the trees: T merged trees of Q / T against one of Q
            q = a.queries // int(v)
            rows.append(report(f"n_trees={v} x Q={q}", gpu_runs(q, seeds, n_trees=int(v)), orc))
        else:
            kw = {a.knob: int(v)}
            R = int(a.rollouts.split(",")[0])
            if R and a.knob != "rollouts_per_leaf":
                kw["rollouts_per_leaf"] = R
            if a.inflight != "1184" and a.knob != "inflight_max":  # a second fixed knob: --inflight N
                kw["inflight_max"] = int(a.inflight)
            rows.append(report(f"{a.knob}={v} Q={a.queries}" + (f" rollouts_per_leaf={R}" if R else ""),
                               gpu_runs(a.queries, seeds, **kw), orc))
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


def cmd_equal_time(a):
    """Which setting gives the best decision in the same wall time?  The query budget of every in-flight cap (and,
    with --rollouts, of every rollouts-per-leaf setting) is what fits `budget_ms` at that setting's own rate
    (measured on a first run); the yardstick is the 1e6-query oracle."""
    seeds = range(a.seeds)
    orc = oracle_values(1_000_000, seeds, a.oracle_cache)
    rows = []
    for R in [int(x) for x in a.rollouts.split(",")]:
        for v in a.values.split(","):
            kw = {"inflight_max": int(v)}
            if R:
                kw["rollouts_per_leaf"] = R
            trial = gpu_runs(200_000, [99], **kw)[0]
            q = int(200_000 * a.budget_ms / trial[2])
            for _ in range(2):  # the rate grows with the tree: refine
                t2 = gpu_runs(q, [98], **kw)[0]
                q = int(q * a.budget_ms / t2[2])
            rows.append(report(f"inflight_max={v} rollouts_per_leaf={R or 64} Q={q} (budget {a.budget_ms} ms)",
                               gpu_runs(q, seeds, **kw), orc))
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


def cmd_cycles(a):
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    import bench
    pomdp = rs78()
    parts, _ = bench.initial_particles(pomdp, bench.N_PARTICLES)
    Q = a.queries
    for infl in [int(x) for x in a.inflight.split(",")]:
        for R in [int(x) for x in a.rollouts.split(",")]:
            solver = pj.POMCPSolver(c=bench.UCB_C, rng=bench.SOLVER_SEED, tree_queries=Q, search_mode="batched",
                                    max_nodes=3 * Q, inflight_max=infl, rollouts_per_leaf=R)
            pl = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
            rows = []
            for it in range(6):
                pl.set_belief_particles(parts)
                pl.reset_counters()
                _, _, V = pl.search(Q)
                ms, _ = pl.phase_ms(abi.PHASE_ROLLOUT)
                cs = pl.counters()
                cy = pl.worker_cycles().astype(np.float64)
                sims = max(cs["simulations"], 1)
                if it:
                    rows.append([ms, cs["rollout_steps"] / sims, cs["tree_steps"] / sims, cs["nodes"], V.max()] + list(cy[:8] / sims))
            r = np.array(rows)
            mean, lo, hi = r.mean(0), r.min(0), r.max(0)
            ms = mean[0]
            print(f"inflight={infl} R={R}: kernel {ms:.2f} ms [{lo[0]:.2f}..{hi[0]:.2f}], sims/s {Q / ms * 1e3:.3e}, "
                  f"steps/s {mean[1] * Q / ms * 1e3:.3e}, depth/sim {mean[2]:.2f}, steps/sim {mean[1]:.0f}, nodes {mean[3]:.0f}, "
                  f"Vbest {mean[4]:.3f} [{lo[4]:.2f}..{hi[4]:.2f}]")
            print(f"   cycles/sim: descent {mean[5]:.0f} insert {mean[6]:.0f} rollout {mean[7]:.0f} backup {mean[8]:.0f} "
                  f"| sum {mean[5:9].sum():.0f} | admitted-life/sim {mean[10]:.0f} | per level {mean[5] / mean[2]:.0f} "
                  f"| rollout per lane-step {mean[7] / (mean[1] / (R or 64)):.0f}")
            if os.environ.get("POMCP_B200_LIB", "").endswith("_ab_probe.so"):  # -DPOMCP_PROBE: dbg[4..7] = descent segments
                lv = mean[2]
                seg = mean[9:13] / lv
                print(f"   per level: wait statistics {seg[0]:.0f} | UCB+argmax {seg[1]:.0f} | pick+publish (insert included) "
                      f"{seg[2]:.0f} | issue child loads + speculative step {seg[3]:.0f} | wait child N/flag "
                      f"{(mean[5] + mean[6]) / lv - seg.sum():.0f}")
            pl.close()


def cmd_sir(a):
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    name = a.problem
    pomdp = rs78() if name == "rs78" else pj.LightDark1D()
    act, o = (5, 0) if name == "rs78" else (2, 3.0)
    rng = np.random.default_rng(0)
    rows = []
    for lg in [int(x) for x in a.log2.split(",")]:
        n = 1 << lg if lg < 64 else lg  # values >= 64 are literal particle counts
        if name == "rs78":
            parts = (np.uint32(0) | np.uint32(3 << 4) | (rng.integers(0, 256, n).astype(np.uint32) << 8))
        else:
            parts = np.zeros(n, dtype=pomdp.state_dtype)
            parts["y"] = rng.normal(2, 3, n)
        S = parts.dtype.itemsize
        for streaming in (False, True):
            if streaming:
                os.environ["POMCP_PF_STREAMING"] = "1"
            else:
                os.environ.pop("POMCP_PF_STREAMING", None)
            pl = pj.solve(pj.POMCPSolver(tree_queries=100, search_mode="batched"), pomdp, belief_mode=abi.BELIEF_EXTERNAL)
            pl.set_belief_particles(parts)
            ts = []
            for it in range(8):
                pl.flush_l2()
                pl.pf_update_sir(act, o if name != "rs78" else (it & 1), seed=it, n_out=n)
                ms, nl = pl.phase_ms(abi.PHASE_PF)
                ts.append(ms)
            t = float(np.median(ts[2:]))
            gbs = n * (2 * S + 16) / t / 1e6
            rows.append({"problem": name, "particles": n, "path": "streaming" if streaming else "auto", "launches": nl,
                         "us": t * 1e3, "alg_GBs": gbs, "frac_of_6540": gbs / 6540.5})
            print(json.dumps(rows[-1]), flush=True)
            pl.close()
    if a.out:
        json.dump(rows, open(a.out, "w"), indent=1)


def cmd_rates(a):
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    cases = [("Tiger", pj.TigerPOMDP(), 100.0, 100000), ("Baby", pj.BabyPOMDP(-5.0, -10.0), 10.0, 100000),
             ("RockSample(7,8)", rs78(), 20.0, 1000000),
             ("RockSample(11,11)", pj.RockSamplePOMDP(11, 11, seed=11011), 20.0, 1000000),
             ("LightDark1D", pj.LightDark1D(), 1.0, 100000)]
    for name, pomdp, c, Q in cases:
        pl = pj.solve(pj.POMCPSolver(c=c, rng=3, tree_queries=Q, search_mode="batched", max_nodes=3 * Q), pomdp,
                      belief_mode=abi.BELIEF_EXTERNAL)
        ms_l = []
        for it in range(4):
            pl.set_belief_initial()
            pl.reset_counters()
            pl.search(Q)
            ms_l.append(pl.phase_ms(abi.PHASE_SEARCH)[0])
            cs = pl.counters()
        ms = float(np.median(ms_l[1:]))
        print(f"{name:18s} Q={Q:8d}: {ms:7.2f} ms, {Q / ms * 1e3:.2e} sims/s, {cs['rollout_steps'] / ms * 1e3:.2e} rollout "
              f"steps/s, depth/sim {cs['tree_steps'] / Q:.1f}, steps/sim {cs['rollout_steps'] / Q:.0f}, nodes {cs['nodes']}")
        pl.close()


def cmd_step(a):
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    import bench
    Q = a.queries
    pomdp = rs78()
    kw = {"n_trees": a.trees} if a.trees > 1 else {}
    solver = pj.POMCPSolver(c=bench.UCB_C, rng=bench.SOLVER_SEED, tree_queries=Q, search_mode="batched",
                            max_nodes=3 * Q * max(a.trees, 1), rollouts_per_leaf=int(a.rollouts.split(",")[0]), **kw)
    policy = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    parts, _ = bench.initial_particles(pomdp, bench.N_PARTICLES)
    policy.set_belief_particles(parts)
    for k in range(a.repeat):
        policy.reset_counters()
        act, N, V = policy.search(Q)
        cs = policy.counters()  # of this decision: ncu profiles the second one (-s 1)
        ms = policy.phase_ms(abi.PHASE_ROLLOUT)[0]
        print(json.dumps({"decision": k, "kernel_ms": ms, "simulations": cs["simulations"], "rollout_steps": cs["rollout_steps"],
                          "tree_steps": cs["tree_steps"], "nodes": cs["nodes"]}), flush=True)
        policy.pf_update_sir(5, k & 1, seed=k, n_out=bench.N_PARTICLES)
    policy.close()


def cmd_c4(a):
    """BASELINE configs[3] on 1 or 2 GPUs (run under torchrun for 2): LightDark1D, 10^6 particles, SIR update +
    plain POMCP (10^5 queries per GPU, root-parallel: the second GPU doubles the simulations behind a decision; the
    particle set is replicated and every rank runs the same SIR update with the same seed - no exchange)."""
    import torch
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    from pomcp_jl_b200.api import nccl_unique_id
    rank, local, world = int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ld = pj.LightDark1D()
    n, Q = 1_000_000, 100_000
    pl = pj.solve(pj.POMCPSolver(c=1.0, rng=4, tree_queries=Q, search_mode="batched", device=local, rank=rank), ld,
                  belief_mode=abi.BELIEF_EXTERNAL)
    if world > 1:
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        pl.comm_init(uid[0], rank, world)
    parts = np.zeros(n, dtype=ld.state_dtype)
    parts["y"] = np.random.default_rng(11).normal(2.0, 3.0, n)
    pl.set_belief_particles(parts)
    rows = []
    for k in range(a.repeat + 3):
        pl.flush_l2()
        pl.sync()
        pl.event_record(0)
        act, N, V = pl.search_rootparallel(Q) if world > 1 else pl.search(Q)
        pl.pf_update_sir(2 if k % 2 == 0 else 0, 3.0 if k % 2 == 0 else 2.0, seed=100 + k, n_out=n)
        pl.event_record(1)
        rows.append((pl.event_elapsed_ms(0, 1), pl.phase_ms(abi.PHASE_SEARCH)[0], pl.phase_ms(abi.PHASE_PF)[0],
                     pl.phase_ms(abi.PHASE_MERGE)[0] if world > 1 else 0.0, int(N.sum())))
    r = np.array(rows[3:])
    t = torch.tensor(r.mean(0), dtype=torch.float64, device="cuda")
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    if rank == 0:
        m = t.cpu().numpy()
        print(json.dumps({"config": "C4 LightDark1D, 1e6 particles SIR + plain POMCP 1e5 queries per GPU", "gpus": world,
                          "decision_ms": m[0], "search_ms": m[1], "sir_us": m[2] * 1e3, "merge_us": m[3] * 1e3,
                          "simulations_behind_a_decision": int(m[4]) + world,
                          "sir_alg_GBs": n * 48 / (m[2] * 1e-3) / 1e9}), flush=True)
    pl.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("cmd", choices=["oracle-values", "knob", "equal-time", "cycles", "sir", "rates", "step", "c4"])
    ap.add_argument("--queries", type=int, default=1_000_000)
    ap.add_argument("--seeds", type=int, default=12)
    ap.add_argument("--knob", default="inflight_max")
    ap.add_argument("--values", default="64,256,1184")
    ap.add_argument("--budget-ms", type=float, default=14.0)
    ap.add_argument("--oracle-cache", default=os.path.join(ROOT, "profiles", "r2_oracle_values_1e6.json"))
    ap.add_argument("--inflight", default="1184")
    ap.add_argument("--rollouts", default="0")
    ap.add_argument("--problem", default="rs78")
    ap.add_argument("--log2", default="12,20,24,26")
    ap.add_argument("--trees", type=int, default=1)
    ap.add_argument("--repeat", type=int, default=2)
    ap.add_argument("--procs", type=int, default=1)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    {"oracle-values": cmd_oracle_values, "knob": cmd_knob, "equal-time": cmd_equal_time, "cycles": cmd_cycles,
     "sir": cmd_sir, "rates": cmd_rates, "step": cmd_step, "c4": cmd_c4}[a.cmd](a)


if __name__ == "__main__":
    main()

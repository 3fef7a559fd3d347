#!/usr/bin/env python
"""bench.py — headline benchmark of the POMCP hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--config c3|c5]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

--config c3 (default; BASELINE.json configs[2], the configuration the metric's target is quoted on):
RockSample(7,8), tree_queries = 10^6 per action, uniform-random rollouts, belief = 2^20 particles maintained by
the SIR particle filter.  One step = one decision: action() [10^6 simulations: descent + rollout + backup + root
arg-max] followed by update() [SIR: propagate, weight, scan, systematic resample of 2^20 particles].  N GPUs:
root-parallel, one independent tree of 10^6 simulations per GPU, root N / N*V merged by one NCCL all-reduce per
decision (weak scaling).

--config c5 (BASELINE.json configs[4]): RockSample(11,11), root-parallel 8 trees x 10^7 simulations per decision,
8 / N trees per GPU searched side by side in one launch, merged on the device and then by the NCCL all-reduce
(strong scaling: the 8 trees are the fixed quantity).

metric = rollout steps/s (generative-model steps executed inside rollouts, counted on the device), whole job;
`tree_sims_per_sec` and `rollouts_per_leaf` are printed beside it in both arms: the batched planner averages
R = 64 rollouts per leaf, the reference algorithm 1, so rollout steps compare like for like only at equal R —
the reference arm therefore runs the CPU restatement at the GPU arm's R (and reports the stock R = 1 run in
`stock_R1`).  `value`: belief already resident in HBM.  `e2e`: the same decision through the public API (action /
update of pomcp_jl_b200) with the particle set in pinned HOST memory, host<->device copies inside the timed region.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

N_PARTICLES = 1 << 20
UCB_C = 20.0
SOLVER_SEED = 20261019
METRIC = "rollout_steps_per_sec"
UNIT = "steps/s"
GPU_ROLLOUTS_PER_LEAF = 64  # the batched planner's default (include/pomcp_b200.h: rollouts_per_leaf)

CONFIGS = {
    "c3": dict(rs_n=7, rs_k=8, rs_seed=7008, tree_queries=1_000_000, trees_total=None, sir=True, scaling="weak",
               name="RockSample(7,8) tree_queries=1e6 per action, random rollouts, SIR belief 2^20 particles "
                    "(BASELINE.json configs[2])"),
    "c5": dict(rs_n=11, rs_k=11, rs_seed=11011, tree_queries=10_000_000, trees_total=8, sir=False, scaling="strong",
               name="RockSample(11,11) root-parallel 8 trees x 1e7 simulations per decision, random rollouts, belief 2^20 "
                    "particles (BASELINE.json configs[4])"),
}
# names the probes import
RS_N, RS_K, RS_SEED, TREE_QUERIES = 7, 8, 7008, 1_000_000


def workload_config(cfg_name, n_gpus):
    c = CONFIGS[cfg_name]
    d = {
        "workload": c["name"], "problem": f"RockSample({c['rs_n']},{c['rs_k']})", "rocks_seed": c["rs_seed"],
        "tree_queries": c["tree_queries"], "particles": N_PARTICLES, "c": UCB_C, "eps": 0.01, "discount": 0.95,
        "rollouts_per_leaf": GPU_ROLLOUTS_PER_LEAF,
        "l2": "flushed between timed steps (256 MiB write); the node pool (>400 MB) also exceeds the 126 MB L2",
    }
    if c["sir"]:
        d["step"] = "action(search 1e6 sims)+update(SIR 2^20)"
        d["parallelism"] = (f"root-parallel x{n_gpus} (one tree per GPU, NCCL all-reduce of root N,N*V)" if n_gpus > 1
                            else "single tree")
    else:
        t = c["trees_total"] // max(n_gpus, 1)
        d["step"] = "action(8 trees x 1e7 sims, merged)"
        d["parallelism"] = (f"root-parallel: {c['trees_total']} trees = {n_gpus} GPU(s) x {t} tree(s) per GPU in one launch, "
                            "merged on the device" + (" and by one NCCL all-reduce of root N,N*V" if n_gpus > 1 else ""))
    return d


def make_pomdp(cfg_name):
    import pomcp_jl_b200 as pj
    c = CONFIGS[cfg_name]
    return pj.RockSamplePOMDP(c["rs_n"], c["rs_k"], seed=c["rs_seed"])


def initial_particles(pomdp, n, pinned=False, seed=RS_SEED):
    rng = np.random.default_rng(seed)
    rocks = rng.integers(0, 1 << pomdp.k, n).astype(np.uint32)
    p = (np.uint32(pomdp.start[0]) | np.uint32(pomdp.start[1] << 4) | (rocks << 8)).astype(np.uint32)
    if pinned:
        import torch
        t = torch.empty(n, dtype=torch.int32).pin_memory()
        arr = t.numpy().view(np.uint32)
        arr[:] = p
        return arr, t  # keep `t` referenced: it owns the pinned allocation
    return p, None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None
        self.t0 = self.t1 = None

    def mark_start(self):
        self.t0 = time.time()

    def mark_stop(self):
        self.t1 = time.time()

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "50", "-i", str(self.gpu)], stdout=subprocess.PIPE,
                                         stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.time(), [c.strip() for c in line.split(",")]))

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        lo = (self.t0 or 0) + 0.02
        hi = (self.t1 or 1e30) + 0.05
        for ts, r in self.rows:
            if not (lo <= ts <= hi):
                continue
            try:
                sm.append(float(r[1]))
                mx.append(float(r[2]))
                for nm, v in zip(names, r[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(nm)
            except Exception:
                continue
        hi = [x for x in sm if x >= 0.5 * max(sm)] if sm else []
        return {"sm_mhz": statistics.median(hi) if hi else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)", float(d.get("sm_max_mhz", 1965.0))
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)", 1965.0


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("LOCAL_RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))


def host_threads():
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: not a property of the box)
    return max(len(os.sched_getaffinity(0)), 1)


# ----------------------------------------------------------------------------- CPU side (oracle/): reference arm + baseline
def oracle_params(cfg_name, rollouts_per_leaf, tree_queries):
    """Solver parameters for the CPU restatement, built WITHOUT touching the CUDA library."""
    from oracle import oracle as O
    from pomcp_jl_b200 import _abi as abi
    return O.default_params(c=UCB_C, seed=SOLVER_SEED, tree_queries=tree_queries, search_mode=abi.MODE_EXACT,
                            belief_mode=abi.BELIEF_EXTERNAL, rollouts_per_leaf=rollouts_per_leaf)


def cpu_trees(cfg_name, n_threads, rollouts_per_leaf, total_queries, slices):
    """`n_threads` host threads, each growing ONE tree of `total_queries` simulations with the sequential
    restatement, in `slices` equal slices (one slice per bench step: a bounded sample that still walks the
    whole 1e6-query tree, so the early shallow-tree phase with its longer rollouts is neither over- nor
    under-represented).  Returns per-slice (seconds, rollout steps, simulations) summed over the threads."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import oracle as O
    pomdp = make_pomdp(cfg_name)
    parts, _ = initial_particles(pomdp, N_PARTICLES)
    planners = []
    for t in range(n_threads):
        sp = oracle_params(cfg_name, rollouts_per_leaf, total_queries)
        sp.seed = SOLVER_SEED + 1000 * t
        op = O.OraclePlanner(pomdp, sp)
        op.set_belief_particles(parts)
        planners.append(op)
    per = max(total_queries // slices, 1)
    out = []
    prev = [dict(rollout_steps=0, simulations=0) for _ in planners]
    with ThreadPoolExecutor(n_threads) as ex:  # ctypes releases the GIL inside orc_search
        for s in range(slices):
            t0 = time.perf_counter()
            rcs = list(ex.map(lambda op: op.search(per)[0], planners))
            dt = time.perf_counter() - t0
            assert all(rc == 0 for rc in rcs), rcs
            cs = [op.counters() for op in planners]
            steps = sum(c["rollout_steps"] - p["rollout_steps"] for c, p in zip(cs, prev))
            sims = sum(c["simulations"] - p["simulations"] for c, p in zip(cs, prev))
            prev = cs
            out.append((dt, steps, sims))
    for op in planners:
        op.close()
    return out


def cpu_sir_seconds(cfg_name):
    from oracle import oracle as O
    pomdp = make_pomdp(cfg_name)
    parts, _ = initial_particles(pomdp, N_PARTICLES)
    pf = O.OraclePlanner(pomdp, oracle_params(cfg_name, 1, 1000))
    pf.set_belief_particles(parts)
    t0 = time.perf_counter()
    assert pf.pf_update_sir(5, 0, seed=1, n_out=N_PARTICLES, mode=1) == 0
    dt = time.perf_counter() - t0
    pf.close()
    return dt


def run_reference(args):
    """The reference's CPU implementation of the path = the sequential C restatement (oracle/; Julia is not
    installed here or on the GPU box, and POMCP.jl targets Julia 0.5).  Rank 0 only.

    Top-level numbers: every host thread grows one independent tree of the FULL per-tree budget at the GPU
    arm's rollouts_per_leaf (equal work per simulation, so rollout steps/s compare like for like); one bench
    step = a 1/(steps + warmup) slice of those trees + one SIR update, the warm-up slices being the shallow
    start of the trees.  `stock_R1`: the reference algorithm as shipped (1 rollout per leaf), one tree x the full
    budget on ONE thread, and the same on all threads."""
    rank, _, _ = dist_env()
    if rank != 0:
        return
    cfg = CONFIGS[args.config]
    Q = cfg["tree_queries"]
    cores = host_threads()
    R = GPU_ROLLOUTS_PER_LEAF
    # per-tree budget of the timed run: the full decision for c3; c5's 1e7-query trees at R = 64 would take ~10 min
    # per thread, so the sample is the first 1e6 queries of each tree (stated in `sample`)
    Q_run = min(Q, 1_000_000)
    slices = args.steps + args.warmup
    sir_s = cpu_sir_seconds(args.config) if cfg["sir"] else 0.0
    sl = cpu_trees(args.config, cores, R, Q_run, slices)
    timed = sl[args.warmup:]
    dt = sum(x[0] for x in timed) + sir_s * len(timed)
    steps_total = sum(x[1] for x in timed)
    sims_total = sum(x[2] for x in timed)
    val = steps_total / dt
    # the algorithm as shipped: 1 rollout per leaf
    one = cpu_trees(args.config, 1, 1, Q_run, 1)[0]
    allt = cpu_trees(args.config, cores, 1, Q_run, 1)[0]
    sample = (f"{cores} host threads x one independent tree of {Q_run} queries each at {R} rollouts per leaf, grown in "
              f"{slices} slices (the {args.steps} timed slices = queries {args.warmup * Q_run // slices}..{Q_run} of every tree)"
              + (f" + one 2^20-particle SIR update per step ({sir_s * 1e3:.0f} ms, one thread)" if cfg["sir"] else "")
              + ("" if Q_run == Q else f"; the decision's trees hold {Q} queries each: bounded sample"))
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": cfg["scaling"],
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.config, args.gpus),
        "tree_sims_per_sec": sims_total / dt, "rollouts_per_leaf": R,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "stock_R1": {
            "rollouts_per_leaf": 1,
            "one_thread": {"value": one[1] / (one[0] + sir_s), "tree_sims_per_sec": one[2] / (one[0] + sir_s),
                           "seconds_per_decision": one[0] + sir_s, "sample": f"one tree x {Q_run} queries + SIR, 1 thread"},
            "all_threads": {"value": allt[1] / (allt[0] + sir_s), "tree_sims_per_sec": allt[2] / (allt[0] + sir_s),
                            "cores": cores, "sample": f"{cores} independent trees x {Q_run} queries, one per thread"},
        },
        "note": "Julia is not installed here or on the GPU box and POMCP.jl targets Julia 0.5: the reference arm is "
                "the sequential C restatement of the reference algorithm (oracle/), not the Julia code.  The reference "
                "uses several cores by running independent episodes in separate processes "
                "(notebooks/Sanity_Checks.ipynb:47,70-77): one independent tree per host thread is that usage.",
    }
    print(json.dumps(line), flush=True)


def cpu_baseline_single_thread(cfg_name):
    """oracle, 1 thread, bounded sample of the same decision (rank 0, N=1 only): the stock algorithm (R = 1) on
    the whole 1e6-query tree, and the GPU arm's R = 64 on the first 2e5 queries of a tree."""
    cfg = CONFIGS[cfg_name]
    Q_run = min(cfg["tree_queries"], 1_000_000)
    sir_s = cpu_sir_seconds(cfg_name) if cfg["sir"] else 0.0
    one = cpu_trees(cfg_name, 1, 1, Q_run, 1)[0]
    eq = cpu_trees(cfg_name, 1, GPU_ROLLOUTS_PER_LEAF, 200_000, 1)[0]
    return {"value": eq[1] / eq[0], "unit": UNIT, "cores": 1, "kind": "port",
            "rollouts_per_leaf": GPU_ROLLOUTS_PER_LEAF, "tree_sims_per_sec": eq[2] / eq[0],
            "sample": f"1 tree, first 200000 of the {cfg['tree_queries']} queries at {GPU_ROLLOUTS_PER_LEAF} rollouts per leaf "
                      f"({eq[0]:.1f} s of CPU; a young tree has longer rollouts, so steps/s is on the high side and sims/s on "
                      "the low side of the full decision's)",
            "stock_R1": {"value": one[1] / (one[0] + sir_s), "tree_sims_per_sec": one[2] / (one[0] + sir_s),
                         "seconds_per_decision": one[0] + sir_s,
                         "sample": f"1 tree x {Q_run} queries at 1 rollout per leaf" + (" + one 2^20-particle SIR update" if cfg["sir"] else "")}}


# ----------------------------------------------------------------------------- GPU arm
def config_latencies():
    """The other BASELINE configs are latency-bound at their sizes: report microseconds."""
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    out = {}

    def med(f, n=12):
        xs = []
        for i in range(n):
            xs.append(f(i))
        return float(statistics.median(xs[2:]))

    # C1: Tiger, 1000 queries, unweighted update
    pl = pj.solve(pj.POMCPSolver(c=100.0, rng=1, tree_queries=1000, search_mode="batched"), pj.TigerPOMDP())

    def c1(i):
        pl.set_belief_initial()
        pl.search(1000)
        return pl.phase_ms(abi.PHASE_SEARCH)[0] * 1e3
    out["c1_tiger_1e3_queries_search_us"] = med(c1)
    pl.close()
    # C2: Baby, 1e5 queries, SIR 1e4 particles
    pl = pj.solve(pj.POMCPSolver(c=10.0, rng=1, tree_queries=100000, search_mode="batched"), pj.BabyPOMDP(-5.0, -10.0),
                  belief_mode=abi.BELIEF_EXTERNAL)
    parts = (np.random.default_rng(2).random(10000) < 0.1).astype(np.uint32)
    pl.set_belief_particles(parts)

    def c2s(i):
        pl.reset_tree()
        pl.search(100000)
        return pl.phase_ms(abi.PHASE_SEARCH)[0] * 1e3

    def c2u(i):
        pl.pf_update_sir(0, i & 1, seed=i, n_out=10000)
        return pl.phase_ms(abi.PHASE_PF)[0] * 1e3
    out["c2_baby_1e5_queries_search_us"] = med(c2s, 6)
    out["c2_baby_sir_1e4_particles_us"] = med(c2u)
    pl.close()
    # C4: LightDark1D, SIR 1e6 particles (16-byte states), 1e5-query search
    ld = pj.LightDark1D()
    pl = pj.solve(pj.POMCPSolver(c=1.0, rng=1, tree_queries=100000, search_mode="batched"), ld, belief_mode=abi.BELIEF_EXTERNAL)
    parts = np.zeros(1_000_000, dtype=ld.state_dtype)
    parts["y"] = np.random.default_rng(11).normal(2.0, 3.0, 1_000_000)

    def c4u(i):
        pl.set_belief_particles(parts)
        pl.flush_l2()
        pl.pf_update_sir(2, 3.0, seed=i, n_out=1_000_000)
        return pl.phase_ms(abi.PHASE_PF)[0] * 1e3

    def c4s(i):
        pl.reset_tree()
        pl.search(100000)
        return pl.phase_ms(abi.PHASE_SEARCH)[0] * 1e3
    us = med(c4u, 8)
    out["c4_lightdark_sir_1e6_particles_us"] = us
    out["c4_lightdark_sir_alg_GBs"] = 1_000_000 * 48 / us / 1e3
    out["c4_lightdark_1e5_queries_search_us"] = med(c4s, 6)
    pl.close()
    # SIR at the sizes the HBM roofline is about (tiled streaming path, sets beyond one co-resident grid): RockSample
    # particles (4-byte states), algorithmic bytes 2 S + 16 per particle over the measured copy bandwidth
    peak = measured_peaks()[0]
    rs = make_pomdp("c3")
    pl = pj.solve(pj.POMCPSolver(tree_queries=100, search_mode="batched"), rs, belief_mode=abi.BELIEF_EXTERNAL)
    for lg in (24, 26):
        n = 1 << lg
        big = (np.uint32(0) | np.uint32(3 << 4) | (np.random.default_rng(lg).integers(0, 256, n).astype(np.uint32) << 8))
        pl.set_belief_particles(big)

        def tiled(i):
            pl.flush_l2()
            pl.pf_update_sir(5, i & 1, seed=i, n_out=n)
            return pl.phase_ms(abi.PHASE_PF)[0] * 1e3
        us = med(tiled, 8)
        gbs = n * 24 / us / 1e3
        out[f"sir_rocksample_2p{lg}_particles"] = {"us": us, "alg_GBs": gbs, "frac_of_hbm_peak": gbs / peak,
                                                   "launches": pl.phase_ms(abi.PHASE_PF)[1]}
    pl.close()
    return out


def narrow_preset(cfg_name, parts, device):
    """The round-1 setting next to the headline: 8 two-warp workers per SM in flight instead of 16
    one-warp workers (same 64 rollouts per leaf).  Same decision, same belief, 6 timed decisions."""
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    cfg = CONFIGS[cfg_name]
    Q = cfg["tree_queries"]
    pomdp = make_pomdp(cfg_name)
    import torch
    n_sm = torch.cuda.get_device_properties(device).multi_processor_count
    solver = pj.POMCPSolver(c=UCB_C, rng=SOLVER_SEED, tree_queries=Q, search_mode="batched", device=device, max_nodes=3 * Q,
                            inflight_max=8 * n_sm)
    pl = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    pl.set_belief_particles(parts)
    ms, steps = [], []
    for k in range(8):
        pl.reset_tree()
        pl.reset_counters()
        pl.flush_l2()
        pl.search(Q)
        ms.append(pl.phase_ms(abi.PHASE_SEARCH)[0])
        steps.append(pl.counters()["rollout_steps"])
    pl.close()
    m, st = statistics.mean(ms[2:]), statistics.mean(steps[2:])
    return {"rollouts_per_leaf": 64, "inflight_max": f"8 per SM ({8 * n_sm})", "ms_per_decision": m,
            "tree_sims_per_sec": Q / m * 1e3, "rollout_steps_per_sec": st / m * 1e3,
            "note": "value estimates sit closer to the sequential algorithm's at equal queries (-1.1 of 10 against -1.6) "
                    "and below the default's at equal time (8.8 against 10.5 in 13 ms): DESIGN.md 4.1"}


def run_gpu(args):
    import torch
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    from pomcp_jl_b200.api import nccl_unique_id

    cfg = CONFIGS[args.config]
    Q = cfg["tree_queries"]
    rank, local_rank, world = dist_env()
    n_gpus = args.gpus
    dist = None
    if world > 1:
        import torch.distributed as dist
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    else:
        torch.cuda.set_device(0)
    device = local_rank if world > 1 else 0
    if cfg["trees_total"]:
        if cfg["trees_total"] % world:
            raise SystemExit(f"--config {args.config}: {cfg['trees_total']} trees do not divide over {world} GPUs")
        T = cfg["trees_total"] // world
    else:
        T = 1

    pomdp = make_pomdp(args.config)
    max_nodes = 3 * Q if T == 1 and Q <= 2_000_000 else int(1.5 * Q * T) + 65536
    solver = pj.POMCPSolver(c=UCB_C, rng=SOLVER_SEED, tree_queries=Q, search_mode="batched", device=device, rank=rank,
                            max_nodes=max_nodes, n_trees=T)
    policy = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    parts, keep = initial_particles(pomdp, N_PARTICLES, pinned=True, seed=cfg["rs_seed"])
    S = parts.dtype.itemsize
    nA = pomdp.n_actions

    if world > 1:
        uid = [nccl_unique_id() if rank == 0 else None]
        dist.broadcast_object_list(uid, src=0)
        policy.comm_init(uid[0], rank, world)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def decide(q=Q):
        if world > 1:
            return policy.search_rootparallel(q)
        return policy.search(q)

    # ---------------- device-resident timing (value)
    policy.set_belief_particles(parts)

    def resident_step(k):
        if cfg["sir"]:
            a, N, V = decide()
            policy.pf_update_sir(5, k & 1, seed=1000 + k, n_out=N_PARTICLES)  # check rock 0, alternating answers
        else:
            policy.reset_tree()  # every action() on a particle belief starts a new RootNode (src/solver.jl:278-281)
            a, N, V = decide()
        return a

    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    nodes_per_search = None
    for k in range(args.warmup):
        if k == args.warmup - 1 and cfg["sir"]:  # untimed: how many belief nodes one decision allocates
            decide()
            nodes_per_search = policy.counters()["nodes"]
        resident_step(k)
    if nodes_per_search is None:
        nodes_per_search = policy.counters()["nodes"]
    policy.reset_counters()
    barrier()
    sampler.mark_start()
    step_ms, pf_ms, search_ms, rollout_ms, merge_ms, pf_launches = [], [], [], [], [], 0
    for k in range(args.steps):
        policy.flush_l2()
        policy.sync()
        policy.event_record(0)
        resident_step(args.warmup + k)
        policy.event_record(1)
        step_ms.append(policy.event_elapsed_ms(0, 1))
        if cfg["sir"]:
            ms, nl = policy.phase_ms(abi.PHASE_PF)
            pf_ms.append(ms)
            pf_launches = nl
        search_ms.append(policy.phase_ms(abi.PHASE_SEARCH)[0])
        rollout_ms.append(policy.phase_ms(abi.PHASE_ROLLOUT)[0])
        if world > 1:
            merge_ms.append(policy.phase_ms(abi.PHASE_MERGE)[0])
    barrier()
    sampler.mark_stop()
    clocks = sampler.stop() if rank == 0 else None
    cs = policy.counters()
    total_ms = sum(step_ms)
    my = torch.tensor([total_ms, float(cs["rollout_steps"]), float(cs["simulations"]), float(cs["kernel_launches"]),
                       float(cs["rollouts"]), sum(rollout_ms), sum(merge_ms)], dtype=torch.float64, device="cuda")
    if dist is not None:
        allv = [torch.zeros_like(my) for _ in range(world)]
        dist.all_gather(allv, my)
        allv = torch.stack(allv).cpu().numpy()
    else:
        allv = my.cpu().numpy()[None, :]
    total_ms_max = float(allv[:, 0].max())
    steps_sum, sims_sum, launches, rollouts_sum = float(allv[:, 1].sum()), float(allv[:, 2].sum()), int(allv[:, 3].sum()), \
        float(allv[:, 4].sum())
    value = steps_sum / (total_ms_max * 1e-3)

    # ---------------- end to end through the public API with host buffers
    sir = pj.SIRParticleFilter(policy, N_PARTICLES, rng=5)
    # host buffers of the particle set: pinned, two of them (the new set is read back into the one the
    # current set does not occupy)
    pinned = [torch.empty(N_PARTICLES, dtype=torch.int32).pin_memory() for _ in range(2)]
    host = [t.numpy().view(np.uint32) for t in pinned]
    host[0][:] = parts
    belief = pj.ParticleCollection(host[0])

    def e2e_step(k, b):
        if world > 1:
            policy.set_belief_particles(b.particles)  # H2D particles
            a = policy.search_rootparallel(Q)[0]     # search + merge, D2H action + root stats
        else:
            a = pj.action(policy, b)                 # H2D particles, search, D2H action + root stats
        if not cfg["sir"]:
            return a, b
        # H2D particles, SIR on device, D2H new particles
        return a, pj.update(sir, b, 5, k & 1, out=host[1] if b.particles.ctypes.data == host[0].ctypes.data else host[0])

    b = belief
    for k in range(max(1, args.warmup // 2)):
        _, b = e2e_step(k, b)
    policy.reset_counters()
    barrier()
    t0 = time.perf_counter()
    for k in range(args.steps):
        _, b = e2e_step(k, b)
    barrier()
    e2e_s = time.perf_counter() - t0
    cs2 = policy.counters()
    my2 = torch.tensor([e2e_s, float(cs2["rollout_steps"]), float(cs2["simulations"])], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx2 = my2.clone()
        dist.all_reduce(mx2, op=dist.ReduceOp.MAX)
        sm2 = my2.clone()
        dist.all_reduce(sm2, op=dist.ReduceOp.SUM)
        e2e_val, e2e_sims = sm2[1].item() / mx2[0].item(), sm2[2].item() / mx2[0].item()
    else:
        e2e_val, e2e_sims = my2[1].item() / e2e_s, my2[2].item() / e2e_s
    h2d = (2 if cfg["sir"] else 1) * N_PARTICLES * S
    d2h = (N_PARTICLES * S if cfg["sir"] else 0) + nA * 16 + 4

    # ---------------- multi-rank correctness, where the driver can see it: the decision every rank holds must be the
    # host-side merge of all the trees' own statistics, and the trees must be distinct complete searches
    merge_check = None
    if world > 1 or T > 1:
        q_chk = min(Q, 200_000)
        policy.set_belief_particles(parts)
        a_m, N_m, V_m = decide(q_chk)
        own = torch.zeros((T, 2, nA), dtype=torch.float64, device="cuda")
        for t in range(T):
            _, rn, n_t, v_t = policy.tree_result(t)
            assert rn == q_chk and n_t.sum() == q_chk - 1, (rn, n_t.sum())
            own[t, 0] = torch.from_numpy(n_t.astype(np.float64))
            own[t, 1] = torch.from_numpy(v_t)
        dec = torch.tensor([float(a_m)] + [float(x) for x in N_m] + [float(x) for x in V_m], dtype=torch.float64, device="cuda")
        if dist is not None:
            owns = [torch.zeros_like(own) for _ in range(world)]
            decs = [torch.zeros_like(dec) for _ in range(world)]
            dist.all_gather(owns, own)
            dist.all_gather(decs, dec)
        else:
            owns, decs = [own], [dec]
        if rank == 0:
            trees = torch.cat(owns).cpu().numpy()          # [world * T, 2, nA]
            decs = torch.stack(decs).cpu().numpy()
            n = trees[:, 0].sum(0)
            v = (trees[:, 0] * trees[:, 1]).sum(0) / np.maximum(n, 1)
            best = 0
            for i in range(1, nA):
                if v[i] >= v[best]:
                    best = i
            same_everywhere = bool((decs == decs[0]).all())
            distinct = len({tuple(x) for x in trees[:, 0]}) == len(trees)
            ok = (same_everywhere and distinct and int(decs[0][0]) == best and np.array_equal(decs[0][1:1 + nA], n)
                  and np.allclose(decs[0][1 + nA:], v, rtol=0, atol=1e-9))
            merge_check = {"ok": bool(ok), "trees": int(len(trees)), "queries_per_tree": q_chk,
                           "same_decision_on_every_rank": same_everywhere, "trees_distinct": bool(distinct),
                           "action": int(decs[0][0]), "host_merge_action": int(best)}
            if not ok:
                raise SystemExit(f"root-parallel merge check FAILED: {merge_check}")

    if rank == 0:
        peak, peak_src, sm_max = measured_peaks()
        roll_s = sum(rollout_ms) * 1e-3
        steps_rank0, sims_rank0 = cs["rollout_steps"], cs["simulations"]
        prof = {}
        pj_path = os.path.join(ROOT, "profiles", "roofline_constants.json")
        if os.path.exists(pj_path):
            prof = json.load(open(pj_path))
        # thread instructions of the timed launches = a * simulations + b * live rollout steps, (a, b) fitted by least
        # squares to single-pass ncu counts of launches that grew different trees (scripts/summarize_profiles.py)
        a_sim, b_step = prof.get("workers_thread_inst_per_sim"), prof.get("workers_thread_inst_per_rollout_step")
        thread_inst = (a_sim * sims_rank0 + b_step * steps_rank0) if (a_sim and b_step) else None
        n_sm = torch.cuda.get_device_properties(device).multi_processor_count
        issue_peak = n_sm * 4 * 32 * sm_max * 1e6  # thread instructions per second: SMs x schedulers x lanes x f_clk
        inst_rate = thread_inst / roll_s if (thread_inst and roll_s > 0) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": cfg["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args.config, n_gpus),
            "tree_sims_per_sec": sims_sum / (total_ms_max * 1e-3),
            "rollouts_per_leaf": policy._sp.rollouts_per_leaf or GPU_ROLLOUTS_PER_LEAF,
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / args.steps * 1e3, "tree_sims_per_sec": e2e_sims},
            "gpu_launches": launches,
            "roofline": {
                "kernel": "k_search_workers (K1-K4: warp-per-path descent, lane-per-trajectory rollouts, backup)",
                "bound": "issue", "unit": "Gthread-inst/s",
                "achieved": inst_rate / 1e9 if inst_rate else None, "peak": issue_peak / 1e9,
                "frac": inst_rate / issue_peak if inst_rate else None,
                "peak_source": f"{n_sm} SMs x 4 schedulers x 32 lanes x {sm_max:.0f} MHz (MEASURED_PEAKS.json sm_max_mhz)",
                "traffic": prof.get("workers_dram_bytes_per_launch"),
                "share_of_step": roll_s * 1e3 / total_ms,
                "launch_ms": statistics.mean(rollout_ms),
                "rollout_steps_per_sec_in_kernel": steps_rank0 / roll_s if roll_s > 0 else None,
                "thread_inst_model": {"per_sim": a_sim, "per_rollout_step": b_step,
                                      "fit": prof.get("workers_inst_fit")},
                "thread_inst_per_rollout_step_all_in": (thread_inst / steps_rank0) if (thread_inst and steps_rank0) else None,
                "warp_efficiency": prof.get("rollout_warp_efficiency"),
                "issue_active_pct_ncu": prof.get("workers_issue_active_pct"),
                "pipes_pct_ncu": prof.get("workers_pipe_pct"),
                "tree": {"avg_descent_depth": cs["tree_steps"] / max(sims_rank0, 1),
                         "nodes_per_sec": nodes_per_search / (statistics.mean(rollout_ms) * 1e-3)
                         if (rollout_ms and nodes_per_search) else None,
                         "rollout_steps_per_sim": steps_rank0 / max(sims_rank0, 1),
                         "dram_bytes_per_sim": (prof["workers_dram_bytes_per_launch"] / TREE_QUERIES)
                         if prof.get("workers_dram_bytes_per_launch") else None,
                         # from the code path: per level two arrival counts (red) and two backup updates (red),
                         # per simulation the root count, the next query id, the spare node id and the child CAS
                         "atomic_ops_per_sim": 4.0 * cs["tree_steps"] / max(sims_rank0, 1) + 4.0},
                "note": "latency-bound by design: the number of simulations in flight is capped (inflight_max) to "
                        "keep the tree no more than that many simulations stale",
            },
            "solver": {"inflight_max": policy._sp.inflight_max or f"16 per SM ({16 * n_sm}), one-warp workers",
                       "inflight_ramp_div": policy._sp.inflight_ramp_div or 32,
                       "rollouts_per_leaf": policy._sp.rollouts_per_leaf or GPU_ROLLOUTS_PER_LEAF, "search_mode": "batched",
                       "trees_per_gpu": T},
            "rollouts_per_sec": rollouts_sum / (total_ms_max * 1e-3),
            "phase_ms_per_step": {"search": statistics.mean(search_ms), "worker_kernel": statistics.mean(rollout_ms),
                                  "sir_update": statistics.mean(pf_ms) if pf_ms else None,
                                  "merge_allreduce": statistics.mean(merge_ms) if merge_ms else None,
                                  "step": statistics.mean(step_ms)},
        }
        if cfg["sir"]:
            pf_s = statistics.mean(pf_ms) * 1e-3
            alg_bytes = N_PARTICLES * (2 * S + 16)
            achieved = alg_bytes / pf_s / 1e9
            line["roofline_pf"] = {
                "kernel": "SIR update K6-K8 (propagate + weight, fixed-point scan, systematic resample)", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": prof.get("pf_dram_bytes_per_update"),
                "algorithmic_bytes": alg_bytes, "bytes_per_particle": 2 * S + 16, "launch_us": pf_s * 1e6,
                "launches_per_update": pf_launches, "peak_source": peak_src, "share_of_step": pf_s * 1e3 / statistics.mean(step_ms),
            }
        if world > 1:
            # what limits multi-GPU efficiency: the slowest rank's worker kernel (its tree grew differently), not the merge
            k = allv[:, 5] / args.steps
            line["per_rank"] = {"worker_kernel_ms": [float(x) for x in k], "worker_kernel_ms_spread": float(k.max() - k.min()),
                                "merge_ms": [float(x) for x in allv[:, 6] / args.steps],
                                "step_ms": [float(x) for x in allv[:, 0] / args.steps]}
        if merge_check is not None:
            line["merge_check"] = merge_check
        if n_gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single_thread(args.config)
    policy.close()
    if rank == 0:
        if n_gpus == 1 and not args.no_extras:
            try:
                line["config_latencies_us"] = config_latencies()
                if cfg["sir"]:
                    line["presets"] = {"narrow": narrow_preset(args.config, np.array(parts), device)}
            except Exception as ex:  # an extra must not take the headline down with it
                line["config_latencies_us"] = {"error": repr(ex)}
        print(json.dumps(line), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c3", choices=sorted(CONFIGS))
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extras", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_gpu(args)


if __name__ == "__main__":
    main()

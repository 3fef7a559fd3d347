#!/usr/bin/env python
"""bench.py — headline benchmark of the POMCP hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[2], the configuration the metric's target is quoted on):
RockSample(7,8), tree_queries = 10^6 per action, uniform-random rollouts, belief = 2^20 particles
maintained by the SIR particle filter.  One step = one decision: action() [10^6 simulations:
descent + rollout + backup + root arg-max] followed by update() [SIR: propagate, weight, scan,
systematic resample of 2^20 particles].  Multi-GPU: root-parallel, one independent tree of 10^6
simulations per GPU, root N / N*V merged by one NCCL all-reduce per decision (weak scaling).

metric = rollout steps/s (generative-model steps executed inside rollouts, counted on the device),
whole job.  `value`: belief already resident in HBM.  `e2e`: the same decision through the public
API (action / update of pomcp_jl_b200) with the particle set in pinned HOST memory, host<->device
copies inside the timed region.
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

RS_N, RS_K, RS_SEED = 7, 8, 7008
TREE_QUERIES = 1_000_000
N_PARTICLES = 1 << 20
UCB_C = 20.0
SOLVER_SEED = 20261019
METRIC = "rollout_steps_per_sec"
UNIT = "steps/s"


def workload_config(n_gpus):
    return {
        "workload": "RockSample(7,8) tree_queries=1e6 per action, random rollouts, SIR belief 2^20 particles "
                    "(BASELINE.json configs[2])",
        "problem": "RockSample(7,8)", "rocks_seed": RS_SEED, "tree_queries": TREE_QUERIES,
        "particles": N_PARTICLES, "c": UCB_C, "eps": 0.01, "discount": 0.95,
        "step": "action(search 1e6 sims)+update(SIR 2^20)",
        "parallelism": f"root-parallel x{n_gpus} (one tree per GPU, NCCL all-reduce of root N,N*V)" if n_gpus > 1
        else "single tree",
        "l2": "flushed between timed steps (256 MiB write); node pool (>400 MB) also exceeds the 126 MB L2",
    }


def initial_particles(pomdp, n, pinned=False):
    rng = np.random.default_rng(RS_SEED)
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


# ----------------------------------------------------------------------------- reference arm (CPU)
def run_reference(args):
    """The reference's CPU implementation of the path = the sequential C restatement (oracle/),
    one independent tree per host thread (root-parallel, merged like the GPU path).  Rank 0 only."""
    rank, _, world = dist_env()
    if rank != 0:
        return
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    from oracle import oracle as O
    pomdp = pj.RockSamplePOMDP(RS_N, RS_K, seed=RS_SEED)
    # all the host threads this process may use (torchrun exports OMP_NUM_THREADS=1: not a property of the box)
    cores = max(len(os.sched_getaffinity(0)), 1)
    sample_q = 100_000  # per tree and per step: bounded sample of the 1e6-query decision
    n_part = N_PARTICLES
    parts, _ = initial_particles(pomdp, n_part)
    solver = pj.POMCPSolver(c=UCB_C, rng=SOLVER_SEED, tree_queries=sample_q, search_mode="exact")
    sp = solver.c_params(abi.BELIEF_EXTERNAL)
    pf = O.OraclePlanner(pomdp, sp)

    def step(k):
        sp.seed = SOLVER_SEED + k
        rc, a, N, V, cs = O.search_rootparallel(pomdp, sp, parts, cores, cores, sample_q)
        assert rc == 0, rc
        pf.set_belief_particles(parts)
        assert pf.pf_update_sir(5, k & 1, seed=k, n_out=n_part, mode=1) == 0
        return cs

    for k in range(args.warmup):
        step(k)
    t0 = time.perf_counter()
    steps_total, sims_total = 0, 0
    for k in range(args.steps):
        cs = step(args.warmup + k)
        steps_total += cs["rollout_steps"]
        sims_total += cs["simulations"]
    dt = time.perf_counter() - t0
    val = steps_total / dt
    sample = (f"{cores} root-parallel trees x {sample_q} queries + one 2^20-particle SIR update per step "
              f"(decision = 1e6 queries on the GPU arm)")
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(args.gpus),
        "tree_sims_per_sec": sims_total / dt,
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "Julia is not installed here or on the GPU box and POMCP.jl targets Julia 0.5: the reference arm is "
                "the sequential C restatement of the reference algorithm (oracle/), not the Julia code",
    }
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------- GPU arm
def cpu_baseline_single_thread():
    """oracle, 1 thread, bounded sample of the same decision (rank 0, N=1 only)."""
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    from oracle import oracle as O
    pomdp = pj.RockSamplePOMDP(RS_N, RS_K, seed=RS_SEED)
    sample_q = TREE_QUERIES
    solver = pj.POMCPSolver(c=UCB_C, rng=SOLVER_SEED, tree_queries=sample_q, search_mode="exact")
    op = O.OraclePlanner(pomdp, solver.c_params(abi.BELIEF_EXTERNAL))
    parts, _ = initial_particles(pomdp, N_PARTICLES)
    t0 = time.perf_counter()
    op.set_belief_particles(parts)
    rc, a, N, V = op.search(sample_q)
    assert rc == 0
    assert op.pf_update_sir(5, 0, seed=1, n_out=N_PARTICLES, mode=1) == 0
    dt = time.perf_counter() - t0
    cs = op.counters()
    return {"value": cs["rollout_steps"] / dt, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": f"1 tree x {sample_q} of the 1e6 queries + one 2^20-particle SIR update, {dt:.1f} s of CPU",
            "tree_sims_per_sec": cs["simulations"] / dt}


def run_gpu(args):
    import torch
    import pomcp_jl_b200 as pj
    from pomcp_jl_b200 import _abi as abi
    from pomcp_jl_b200.api import nccl_unique_id

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

    pomdp = pj.RockSamplePOMDP(RS_N, RS_K, seed=RS_SEED)
    solver = pj.POMCPSolver(c=UCB_C, rng=SOLVER_SEED, tree_queries=TREE_QUERIES, search_mode="batched",
                            device=device, rank=rank, max_nodes=3 * TREE_QUERIES)
    policy = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    parts, keep = initial_particles(pomdp, N_PARTICLES, pinned=True)
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

    def decide():
        if world > 1:
            return policy.search_rootparallel(TREE_QUERIES)
        return policy.search(TREE_QUERIES)

    # ---------------- device-resident timing (value)
    policy.set_belief_particles(parts)
    pf_seed = 0

    def resident_step(k):
        a, N, V = decide()
        policy.pf_update_sir(5, k & 1, seed=1000 + k, n_out=N_PARTICLES)  # check rock 0, alternating answers
        return a

    sampler = ClockSampler(device)
    if rank == 0:
        sampler.start()
    nodes_per_search = None
    for k in range(args.warmup):
        if k == args.warmup - 1:  # untimed: how many belief nodes one decision allocates
            decide()
            nodes_per_search = policy.counters()["nodes"]
        resident_step(k)
    policy.reset_counters()
    barrier()
    sampler.mark_start()
    step_ms, pf_ms, search_ms, rollout_ms, pf_launches = [], [], [], [], 0
    for k in range(args.steps):
        policy.flush_l2()
        policy.sync()
        policy.event_record(0)
        resident_step(args.warmup + k)
        policy.event_record(1)
        step_ms.append(policy.event_elapsed_ms(0, 1))
        ms, nl = policy.phase_ms(abi.PHASE_PF)
        pf_ms.append(ms)
        pf_launches = nl
        search_ms.append(policy.phase_ms(abi.PHASE_SEARCH)[0])
        rollout_ms.append(policy.phase_ms(abi.PHASE_ROLLOUT)[0])
    barrier()
    sampler.mark_stop()
    clocks = sampler.stop() if rank == 0 else None
    cs = policy.counters()
    total_ms = sum(step_ms)
    my = torch.tensor([total_ms, float(cs["rollout_steps"]), float(cs["simulations"]), float(cs["kernel_launches"])],
                      dtype=torch.float64, device="cuda")
    if dist is not None:
        mx = my.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = my.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        total_ms_max, steps_sum, sims_sum, launches = mx[0].item(), sm[1].item(), sm[2].item(), int(sm[3].item())
    else:
        total_ms_max, steps_sum, sims_sum, launches = total_ms, my[1].item(), my[2].item(), int(my[3].item())
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
            policy.set_belief_particles(b.particles)
            a = policy.search_rootparallel(TREE_QUERIES)[0]
        else:
            a = pj.action(policy, b)                 # H2D particles, search, D2H action + root stats
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
    my2 = torch.tensor([e2e_s, float(cs2["rollout_steps"])], dtype=torch.float64, device="cuda")
    if dist is not None:
        mx2 = my2.clone()
        dist.all_reduce(mx2, op=dist.ReduceOp.MAX)
        sm2 = my2.clone()
        dist.all_reduce(sm2, op=dist.ReduceOp.SUM)
        e2e_val = sm2[1].item() / mx2[0].item()
    else:
        e2e_val = my2[1].item() / e2e_s
    h2d = 2 * N_PARTICLES * S
    d2h = N_PARTICLES * S + nA * 16 + 4

    if rank == 0:
        peak, peak_src, sm_max = measured_peaks()
        pf_s = statistics.mean(pf_ms) * 1e-3
        alg_bytes = N_PARTICLES * (2 * S + 16)
        achieved = alg_bytes / pf_s / 1e9
        roll_s = sum(rollout_ms) * 1e-3
        steps_rank0 = cs["rollout_steps"]
        prof = {}
        pj_path = os.path.join(ROOT, "profiles", "roofline_constants.json")
        if os.path.exists(pj_path):
            prof = json.load(open(pj_path))
        # thread instructions of the timed launches = a * simulations + b * live rollout steps, (a, b) fitted to
        # single-pass ncu counts of launches that grew different trees (scripts/summarize_profiles.py)
        n_search = max(len(rollout_ms), 1)
        if prof.get("workers_thread_inst_per_sim") and prof.get("workers_thread_inst_per_rollout_step"):
            thread_inst = (prof["workers_thread_inst_per_sim"] * cs["simulations"] +
                           prof["workers_thread_inst_per_rollout_step"] * steps_rank0)
        elif prof.get("workers_thread_inst_per_launch"):
            thread_inst = prof["workers_thread_inst_per_launch"] * n_search
        else:
            thread_inst = None
        issue_frac = thread_inst / roll_s / (148 * 4 * 32 * sm_max * 1e6) if (thread_inst and roll_s > 0) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n_gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms_max / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(n_gpus),
            "tree_sims_per_sec": sims_sum / (total_ms_max * 1e-3),
            "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_s / args.steps * 1e3,
                    "tree_sims_per_sec": TREE_QUERIES * n_gpus * args.steps / e2e_s},
            "gpu_launches": launches,
            "roofline": {
                "kernel": "SIR update K6-K8: k_pf_fused (one cooperative launch: propagate+weight, fixed-point scan, "
                          "systematic resample)" if pf_launches == 1 else "SIR update K6-K8: k_pf_max + k_pf_scan + "
                          "k_pf_finalize + k_pf_emit", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": prof.get("pf_dram_bytes_per_update"),
                "algorithmic_bytes": alg_bytes, "bytes_per_particle": 2 * S + 16, "launch_us": pf_s * 1e6,
                "launches_per_update": pf_launches, "peak_source": peak_src,
            },
            "roofline_rollout": {
                "kernel": "k_search_workers (K1-K4: warp-per-path descent, lane-per-trajectory rollouts, backup)",
                "bound": "issue",
                "achieved": steps_rank0 / roll_s if roll_s > 0 else None, "unit": "rollout steps/s inside the kernel",
                "share_of_step": roll_s * 1e3 / total_ms,
                "thread_inst_per_step": (thread_inst / steps_rank0) if (thread_inst and steps_rank0) else None,
                "thread_inst_model": {"per_sim": prof.get("workers_thread_inst_per_sim"),
                                      "per_rollout_step": prof.get("workers_thread_inst_per_rollout_step")},
                # frac = thread instructions per launch (ncu, one pass; about constant from tree to tree, unlike
                # the number of live rollout steps) / measured kernel time / (SMs x 4 schedulers x 32 lanes x f_clk):
                # issue-slot utilisation x active lanes.  peak = achieved / frac.
                "frac": issue_frac,
                "peak": (steps_rank0 / roll_s / issue_frac) if (issue_frac and roll_s > 0) else None,
                "warp_efficiency": prof.get("rollout_warp_efficiency"),
                "tree": {"avg_descent_depth": cs["tree_steps"] / max(cs["simulations"], 1),
                         "nodes_per_sec": nodes_per_search / (statistics.mean(rollout_ms) * 1e-3)
                         if (rollout_ms and nodes_per_search) else None,
                         "rollout_steps_per_sim": cs["rollout_steps"] / max(cs["simulations"], 1),
                         "dram_bytes_per_sim": (prof["workers_dram_bytes_per_launch"] / TREE_QUERIES)
                         if prof.get("workers_dram_bytes_per_launch") else None,
                         # from the code path: per level two arrival counts (red) and two backup updates (red),
                         # per simulation the root count, the next query id, the spare node id and the child CAS
                         "atomic_ops_per_sim": 4.0 * cs["tree_steps"] / max(cs["simulations"], 1) + 4.0},
                "issue_active_pct": prof.get("workers_issue_active_pct"),
                "note": "latency-bound by design: the number of simulations in flight is capped (inflight_max) to "
                        "keep the tree no more than that many simulations stale",
            },
            "solver": {"inflight_max": policy._sp.inflight_max or "8 per SM (1184)", "inflight_ramp_div": policy._sp.inflight_ramp_div
                       or 32, "rollouts_per_leaf": policy._sp.rollouts_per_leaf or 64, "search_mode": "batched"},
            "rollouts_per_sec": cs["rollouts"] * n_gpus / (total_ms_max * 1e-3),
            "phase_ms_per_step": {"search": statistics.mean(search_ms), "worker_kernel": statistics.mean(rollout_ms),
                                  "sir_update": statistics.mean(pf_ms), "step": statistics.mean(step_ms)},
        }
        if n_gpus == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline_single_thread()
        print(json.dumps(line), flush=True)
    policy.close()
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3
        run_gpu(args)


if __name__ == "__main__":
    main()

"""Developer probe (GPU box): where a batched worker's cycles go (pomcp_worker_cycles), per setting.

    python scripts/cycle_probe.py [Q] [inflight,...] [rollouts,...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
import bench

Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
inflights = [int(x) for x in sys.argv[2].split(",")] if len(sys.argv) > 2 else [1024]
rolls = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [32]
pomdp = pj.RockSamplePOMDP(bench.RS_N, bench.RS_K, seed=bench.RS_SEED)
parts, _ = bench.initial_particles(pomdp, bench.N_PARTICLES)
for infl in inflights:
    for R in rolls:
        solver = pj.POMCPSolver(c=bench.UCB_C, rng=bench.SOLVER_SEED, tree_queries=Q, search_mode="batched",
                                max_nodes=3 * Q, inflight_max=infl, rollouts_per_leaf=R)
        pl = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
        rows = []
        for it in range(6):
            pl.set_belief_particles(parts)
            pl.reset_counters()
            a, N, V = pl.search(Q)
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
              f"| rollout per lane-step {mean[7] / (mean[1] / R):.0f}")
        if os.environ.get("POMCP_B200_LIB", "").endswith("_ab_probe.so"):  # -DPOMCP_PROBE: dbg[4..7] = descent segments
            lv = mean[2]
            seg = mean[9:13] / lv
            print(f"   per level: wait statistics {seg[0]:.0f} | UCB+argmax {seg[1]:.0f} | pick+publish (insert included) {seg[2]:.0f} "
                  f"| issue child loads + speculative step {seg[3]:.0f} | wait child N/flag {(mean[5] + mean[6]) / lv - seg.sum():.0f}")
        pl.close()

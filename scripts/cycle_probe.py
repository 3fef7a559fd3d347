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
        for it in range(3):
            pl.set_belief_particles(parts)
            pl.reset_counters()
            a, N, V = pl.search(Q)
            ms, _ = pl.phase_ms(abi.PHASE_ROLLOUT)
            cs = pl.counters()
            cy = pl.worker_cycles().astype(np.float64)
        sims = max(cs["simulations"], 1)
        per = cy[:4] / sims
        print(f"inflight={infl} R={R}: kernel {ms:.2f} ms, sims/s {Q / ms * 1e3:.3e}, steps/s {cs['rollout_steps'] / ms * 1e3:.3e}, "
              f"depth/sim {cs['tree_steps'] / sims:.2f}, steps/sim {cs['rollout_steps'] / sims:.0f}, nodes {cs['nodes']}, a={a} Vbest={V.max():.3f}")
        print(f"   cycles/sim: descent {per[0]:.0f} insert {per[1]:.0f} rollout {per[2]:.0f} backup {per[3]:.0f} "
              f"| sum {per.sum():.0f} | admitted-life/sim {cy[5] / sims:.0f} admission-wait total {cy[4]:.3e} | "
              f"per level {per[0] / max(cs['tree_steps'] / sims, 1):.0f}")
        pl.close()

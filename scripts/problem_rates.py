"""Developer probe (GPU box): batched search rates of the four concrete problems with the library defaults."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

cases = [("Tiger", pj.TigerPOMDP(), 100.0, 100000), ("Baby", pj.BabyPOMDP(-5.0, -10.0), 10.0, 100000),
         ("RockSample(7,8)", pj.RockSamplePOMDP(7, 8, seed=7008), 20.0, 1000000),
         ("RockSample(11,11)", pj.RockSamplePOMDP(11, 11, seed=11011), 20.0, 1000000),
         ("LightDark1D", pj.LightDark1D(), 1.0, 100000)]
for name, pomdp, c, Q in cases:
    pl = pj.solve(pj.POMCPSolver(c=c, rng=3, tree_queries=Q, search_mode="batched", max_nodes=3 * Q), pomdp,
                  belief_mode=abi.BELIEF_EXTERNAL)
    ms_l = []
    for it in range(4):
        pl.set_belief_initial()
        pl.reset_counters()
        a, N, V = pl.search(Q)
        ms_l.append(pl.phase_ms(abi.PHASE_SEARCH)[0])
        cs = pl.counters()
    ms = float(np.median(ms_l[1:]))
    print(f"{name:18s} Q={Q:8d}: {ms:7.2f} ms, {Q / ms * 1e3:.2e} sims/s, {cs['rollout_steps'] / ms * 1e3:.2e} rollout steps/s, "
          f"depth/sim {cs['tree_steps'] / Q:.1f}, steps/sim {cs['rollout_steps'] / Q:.0f}, nodes {cs['nodes']}")
    pl.close()

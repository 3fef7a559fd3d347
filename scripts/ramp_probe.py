"""Developer probe (GPU box): admission ramp (inflight_ramp_div) against fidelity and decision time.
RockSample(7,8), c = 20, initial belief, defaults otherwise; oracle = sequential restatement, same seeds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
from oracle import oracle as O

pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
divs = [int(x) for x in sys.argv[1].split(",")] if len(sys.argv) > 1 else [32, 16, 8, 4]
for Q, seeds in ((10000, range(16)), (100000, range(12)), (1000000, range(6))):
    orc = []
    for seed in seeds:
        solver = pj.POMCPSolver(c=20.0, rng=seed, tree_queries=Q, search_mode="exact")
        op = O.OraclePlanner(pomdp, solver.c_params(abi.BELIEF_UNWEIGHTED))
        op.set_belief_initial()
        rc, a, N, V = op.search(Q)
        orc.append((a, V))
        op.close()
    vo = np.array([v[a] for a, v in orc])
    print(f"Q={Q}: oracle best-arm V {vo.mean():.2f} +- {vo.std() / np.sqrt(len(vo)):.2f}", flush=True)
    for div in divs:
        ag, vg, ms_l = [], [], []
        for k, seed in enumerate(seeds):
            solver = pj.POMCPSolver(c=20.0, rng=seed, tree_queries=Q, search_mode="batched", inflight_ramp_div=div,
                                    max_nodes=4 * Q)
            gp = pj.solve(solver, pomdp)
            gp.set_belief_initial()
            a, N, V = gp.search(Q)
            a, N, V = gp.search(Q) if False else (a, N, V)
            ms_l.append(gp.phase_ms(abi.PHASE_SEARCH)[0])
            gp.close()
            ag.append(a == orc[k][0])
            vg.append(V[a])
        print(f"  ramp_div {div:3d}: action = oracle's {sum(ag)}/{len(ag)}, best-arm V {np.mean(vg):.2f} +- "
              f"{np.std(vg) / np.sqrt(len(vg)):.2f}, decision {np.mean(ms_l):.2f} ms", flush=True)

"""Developer probe (GPU box): the fidelity / throughput table of DESIGN.md §4.1 — batched search against
the sequential oracle on RockSample(7,8), c = 20, per number of simulations in flight."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
from oracle import oracle as O

pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)


def oracle_runs(Q, seeds):
    out = []
    for seed in seeds:
        solver = pj.POMCPSolver(c=20.0, rng=seed, tree_queries=Q, search_mode="exact")
        op = O.OraclePlanner(pomdp, solver.c_params(abi.BELIEF_UNWEIGHTED))
        op.set_belief_initial()
        rc, a, N, V = op.search(Q)
        out.append((a, V))
        op.close()
    return out


for Q, seeds in ((100000, range(12)), (1000000, range(4))):
    orc = oracle_runs(Q, seeds)
    vo = np.array([v[a] for a, v in orc])
    print(f"Q={Q}: oracle best-arm V {vo.mean():.2f} +- {vo.std():.2f}, actions {[a for a, _ in orc]}")
    for infl in (64, 256, 1024, 4096):
        ag, vg, ms_l, st_l = [], [], [], []
        for k, seed in enumerate(seeds):
            solver = pj.POMCPSolver(c=20.0, rng=seed, tree_queries=Q, search_mode="batched", inflight_max=infl,
                                    max_nodes=4 * Q)
            gp = pj.solve(solver, pomdp)
            gp.set_belief_initial()
            a, N, V = gp.search(Q)
            ms_l.append(gp.phase_ms(abi.PHASE_SEARCH)[0])
            st_l.append(gp.counters()["rollout_steps"])
            gp.close()
            ag.append(a == orc[k][0])
            vg.append(V[a])
        ms = np.mean(ms_l)
        print(f"  in flight {infl:5d}: action = oracle's {sum(ag)}/{len(ag)}, best-arm V {np.mean(vg):.2f} +- {np.std(vg):.2f}, "
              f"sims/s {Q / ms * 1e3:.2e}, rollout steps/s {np.mean(st_l) / ms * 1e3:.2e}")

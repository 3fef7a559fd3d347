"""Developer probe (GPU box): batched-vs-sequential search statistics across seeds and wave settings."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
from oracle import oracle as O

def run(name, pomdp, c, Q, seeds, **kw):
    rows = []
    for seed in seeds:
        solver = pj.POMCPSolver(c=c, rng=seed, tree_queries=Q, search_mode="batched", **kw)
        gp = pj.solve(solver, pomdp)
        gp.set_belief_initial()
        a_g, N_g, V_g = gp.search(Q)
        ms = gp.phase_ms(abi.PHASE_SEARCH)[0]
        cs = gp.counters()
        gp.close()
        op = O.OraclePlanner(pomdp, solver.c_params(abi.BELIEF_UNWEIGHTED))
        op.set_belief_initial()
        rc, a_o, N_o, V_o = op.search(Q)
        op.close()
        rows.append((a_g, a_o, N_g, N_o, V_g, V_o, ms, cs["rollout_steps"]))
    agree = np.mean([r[0] == r[1] for r in rows])
    Vg = np.array([r[4] for r in rows]); Vo = np.array([r[5] for r in rows])
    Ng = np.array([r[2] for r in rows]); No = np.array([r[3] for r in rows])
    msm = np.mean([r[6] for r in rows]); st = np.mean([r[7] for r in rows])
    print(f"{name} Q={Q} {kw} agree={agree:.2f} search_ms={msm:.2f} sims/s={Q/msm*1e3:.3e} rollout_steps/s={st/msm*1e3:.3e}")
    print("   V_gpu mean", np.round(Vg.mean(0), 2), " sd", np.round(Vg.std(0), 2))
    print("   V_orc mean", np.round(Vo.mean(0), 2), " sd", np.round(Vo.std(0), 2))
    print("   N_gpu mean", np.round(Ng.mean(0)).astype(int))
    print("   N_orc mean", np.round(No.mean(0)).astype(int))
    best_o = [int(np.argmax(r[5])) for r in rows]
    print("   oracle best-action histogram", np.bincount(best_o, minlength=pomdp.n_actions),
          " gpu", np.bincount([r[0] for r in rows], minlength=pomdp.n_actions))

if __name__ == "__main__":
    seeds = range(12)
    for kw in ({"inflight_max": 64}, {"inflight_max": 256}, {"inflight_max": 1024}, {"inflight_max": 4096},
               {"inflight_max": 256, "rollouts_per_leaf": 1}, {"inflight_max": 1024, "inflight_ramp_div": 128},
               {"inflight_max": 4096, "inflight_ramp_div": 128}):
        run("tiger", pj.TigerPOMDP(), 1000.0, 20000, seeds, **kw)
        run("baby", pj.BabyPOMDP(-5., -10.), 10.0, 50000, seeds, **kw)
        run("rs78", pj.RockSamplePOMDP(7, 8, seed=7008), 20.0, 100000, seeds, **kw)
    for kw in ({"inflight_max": 256}, {"inflight_max": 1024}, {"inflight_max": 4096}):
        run("rs78-1M", pj.RockSamplePOMDP(7, 8, seed=7008), 20.0, 1000000, range(3), **kw)

"""Developer probe: ramp / in-flight settings vs the sequential oracle on RockSample(7,8)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
from oracle import oracle as O

rs = pj.RockSamplePOMDP(7, 8, seed=7008)
orc = {}
def oracle_stats(Q, seeds):
    if (Q, len(seeds)) not in orc:
        rows = []
        for seed in seeds:
            op = O.OraclePlanner(rs, O.default_params(c=20.0, seed=seed))
            op.set_belief_initial(); rc, a, N, V = op.search(Q); op.close(); rows.append((a, V.max()))
        orc[(Q, len(seeds))] = rows
    return orc[(Q, len(seeds))]

def run(Q, seeds, **kw):
    ref = oracle_stats(Q, seeds)
    out = []
    for seed in seeds:
        gp = pj.solve(pj.POMCPSolver(c=20.0, rng=seed, tree_queries=Q, search_mode="batched", **kw), rs)
        gp.set_belief_initial(); a, N, V = gp.search(Q); ms = gp.phase_ms(abi.PHASE_SEARCH)[0]; cs = gp.counters(); gp.close()
        out.append((a, V.max(), ms, cs["rollout_steps"]))
    agree = np.mean([o[0] == r[0] for o, r in zip(out, ref)])
    print(f"Q={Q} {kw}: agree={agree:.2f} Vbest gpu={np.mean([o[1] for o in out]):.2f}±{np.std([o[1] for o in out]):.2f} "
          f"oracle={np.mean([r[1] for r in ref]):.2f}±{np.std([r[1] for r in ref]):.2f} ms={np.mean([o[2] for o in out]):.1f} "
          f"sims/s={Q/np.mean([o[2] for o in out])*1e3:.3e} steps/s={np.mean([o[3] for o in out])/np.mean([o[2] for o in out])*1e3:.3e}", flush=True)

for kw in ({}, {"inflight_ramp_div": 16}, {"inflight_ramp_div": 8}, {"inflight_ramp_div": 16, "inflight_max": 2048}, {"inflight_max": 512}):
    run(100000, range(8), **kw)
for kw in ({}, {"inflight_ramp_div": 16}, {"inflight_ramp_div": 8}, {"inflight_ramp_div": 16, "inflight_max": 2048}, {"inflight_max": 512}, {"inflight_max": 2048}):
    run(1000000, range(4), **kw)

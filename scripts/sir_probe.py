import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
rs = pj.RockSamplePOMDP(7, 8, seed=7008)
gp = pj.solve(pj.POMCPSolver(c=20.0, rng=1, tree_queries=1000), rs, belief_mode=abi.BELIEF_EXTERNAL)
rng = np.random.default_rng(0)
sizes = [int(x) for x in sys.argv[1:]] or [1 << 20, 1 << 16, 32]
for n in sizes:
    parts = (np.uint32(0) | np.uint32(3 << 4) | (rng.integers(0, 256, n).astype(np.uint32) << 8))
    ts = []
    for it in range(7):
        gp.set_belief_particles(parts)
        gp.pf_update_sir(5, it & 1, seed=it, n_out=n)
        ts.append(gp.phase_ms(abi.PHASE_PF)[0] * 1e3)
    best = min(ts[2:])
    print(os.environ.get("POMCP_PF_STREAMING", "fused"), n, "us:", np.round(ts[2:], 1), "alg GB/s at best: %.0f" % (n * 24 / best / 1e3), flush=True)
gp.close()

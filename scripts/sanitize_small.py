"""Small end-to-end run for compute-sanitizer (memcheck / racecheck): every kernel once, tiny sizes."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
rng = np.random.default_rng(0)
for pomdp, c in ((pj.RockSamplePOMDP(4, 4, seed=44), 20.0), (pj.LightDark1D(), 1.0), (pj.BabyPOMDP(), 10.0)):
    for mode in ("exact", "batched"):
        pl = pj.solve(pj.POMCPSolver(c=c, rng=3, tree_queries=400, search_mode=mode, inflight_max=64), pomdp)
        pl.set_belief_initial()
        a, N, V = pl.search(400)
        if not pomdp.continuous_obs:
            for o in range(pomdp.n_observations):
                if pl.node_stats([(a, o)])[0] > 0:
                    pl.update_unweighted(a, o)
                    pl.search(300)
                    break
        pl.close()
    pl = pj.solve(pj.POMCPSolver(c=c, rng=3, tree_queries=200), pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    if isinstance(pomdp, pj.LightDark1D):
        s = np.zeros(5000, dtype=pomdp.state_dtype); s["y"] = rng.normal(2, 3, 5000); o = 3.0
    elif isinstance(pomdp, pj.RockSamplePOMDP):
        s = (rng.integers(0, 16, 5000).astype(np.uint32) << 8); o = 2
    else:
        s = rng.integers(0, 2, 5000).astype(np.uint32); o = 1
    pl.set_belief_particles(s)
    pl.search(200)
    pl.pf_update_sir(0, o, seed=1, n_out=5000)
    os.environ["POMCP_PF_STREAMING"] = "1"
    pl2 = pj.solve(pj.POMCPSolver(c=c, rng=3, tree_queries=200), pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    pl2.set_belief_particles(s); pl2.pf_update_sir(0, o, seed=1, n_out=5000)
    assert pl.belief_particles().tobytes() == pl2.belief_particles().tobytes()
    del os.environ["POMCP_PF_STREAMING"]
    pl.rollout_batch(s[:1000], 20, 5); pl.pf_weights(s[:1000], 0, o, 7)
    pl.resample_systematic(rng.random(3000), 0.3, 3000, 0); pl.resample_systematic(rng.random(300), 0.3, 300, 1)
    pl.close(); pl2.close()
print("sanitize run ok")

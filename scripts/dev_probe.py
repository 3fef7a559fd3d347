"""Developer probe (GPU box): timings of the batched search and the SIR update at bench sizes."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

def probe_search(pomdp, c, Q, n_part, **kw):
    solver = pj.POMCPSolver(c=c, rng=1, tree_queries=Q, search_mode="batched", **kw)
    gp = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    rng = np.random.default_rng(0)
    if isinstance(pomdp, pj.RockSamplePOMDP):
        parts = (np.uint32(pomdp.start[0]) | np.uint32(pomdp.start[1] << 4) | (rng.integers(0, 1 << pomdp.k, n_part).astype(np.uint32) << 8))
    elif isinstance(pomdp, pj.LightDark1D):
        parts = np.zeros(n_part, dtype=pomdp.state_dtype); parts["y"] = rng.normal(2, 3, n_part)
    else:
        parts = (rng.random(n_part) < 0.1).astype(np.uint32)
    for it in range(3):
        gp.set_belief_particles(parts)
        gp.reset_counters()
        t = time.time(); a, N, V = gp.search(Q); dt = time.time() - t
        cs = gp.counters()
        ms_s, ls = gp.phase_ms(abi.PHASE_SEARCH); ms_r, lr = gp.phase_ms(abi.PHASE_ROLLOUT)
        print(f"{type(pomdp).__name__} Q={Q} it={it} wall={dt*1e3:.2f}ms search={ms_s:.2f}ms worker-kernel={ms_r:.2f}ms waves={cs['waves']} "
              f"rollout_steps={cs['rollout_steps']} steps/s={cs['rollout_steps']/(ms_s*1e-3):.3e} sims/s={Q/(ms_s*1e-3):.3e} "
              f"nodes={cs['nodes']} maxdepth={cs['max_depth_seen']} a={a} Nmax={N.max()} V={np.round(V,2)}")
    return gp, parts

def probe_sir(gp, pomdp, parts, a, o):
    for it in range(3):
        gp.set_belief_particles(parts)
        t = time.time(); gp.pf_update_sir(a, o, seed=it, n_out=len(parts)); dt = time.time() - t
        ms, l = gp.phase_ms(abi.PHASE_PF)
        S = parts.dtype.itemsize
        print(f"SIR {type(pomdp).__name__} n={len(parts)} wall={dt*1e3:.2f}ms kernels={ms*1e3:.1f}us  algGB/s={(len(parts)*(2*S+16))/(ms*1e-3)/1e9:.1f}")

if __name__ == "__main__":
    Q = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
    rs = pj.RockSamplePOMDP(7, 8, seed=7008)
    gp, parts = probe_search(rs, 20.0, Q, 1 << 20)
    probe_sir(gp, rs, parts, 5, 0)
    gp.close()
    ld = pj.LightDark1D()
    gp, parts = probe_search(ld, 1.0, 100000, 1000000)
    probe_sir(gp, ld, parts, 2, 3.0)
    gp.close()
    gp, parts = probe_search(pj.BabyPOMDP(), 10.0, 100000, 10000)
    probe_sir(gp, pj.BabyPOMDP(), parts, 0, 1)
    gp.close()
    for im in (256, 4096):
        gp, parts = probe_search(rs, 20.0, Q, 1 << 20, inflight_max=im)
        gp.close()

"""Developer probe (GPU box): SIR update (K6-K8) time and algorithmic GB/s per particle count and path.

    python scripts/sir_probe.py [problem=rs78|lightdark] [log2 sizes ...]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

name = sys.argv[1] if len(sys.argv) > 1 else "rs78"
sizes = [int(x) for x in sys.argv[2:]] or [20, 22, 24, 26]
pomdp = pj.RockSamplePOMDP(7, 8, seed=7008) if name == "rs78" else pj.LightDark1D()
a, o = (5, 0) if name == "rs78" else (2, 3.0)
rng = np.random.default_rng(0)
for lg in sizes:
    n = 1 << lg
    if name == "rs78":
        parts = (np.uint32(0) | np.uint32(3 << 4) | (rng.integers(0, 256, n).astype(np.uint32) << 8))
    else:
        parts = np.zeros(n, dtype=pomdp.state_dtype)
        parts["y"] = rng.normal(2, 3, n)
    S = parts.dtype.itemsize
    for streaming in (False, True):
        if streaming:
            os.environ["POMCP_PF_STREAMING"] = "1"
        else:
            os.environ.pop("POMCP_PF_STREAMING", None)
        pl = pj.solve(pj.POMCPSolver(tree_queries=100, search_mode="batched"), pomdp, belief_mode=abi.BELIEF_EXTERNAL)
        pl.set_belief_particles(parts)
        ts = []
        for it in range(6):
            pl.flush_l2()
            pl.pf_update_sir(a, o if name != "rs78" else (it & 1), seed=it, n_out=n)
            ms, nl = pl.phase_ms(abi.PHASE_PF)
            ts.append(ms)
        t = np.median(ts[1:])
        print(f"{name} n=2^{lg} {'streaming' if streaming else 'auto     '} launches={nl} {t * 1e3:9.1f} us  "
              f"alg {n * (2 * S + 16) / t / 1e6:8.1f} GB/s  (frac of 6540: {n * (2 * S + 16) / t / 1e6 / 6540.5:.3f})")
        pl.close()

"""Developer probe (GPU box, probe build: scripts/build_variant.sh probe -DPOMCP_PROBE; POMCP_B200_LIB=scripts/_ab_probe.so):
where the one-barrier SIR kernel spends its time - per-CTA %globaltimer at start / barrier arrival / barrier exit / end."""
import ctypes as C
import os
import sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
from pomcp_jl_b200._lib import lib

n = 1 << int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
rng = np.random.default_rng(0)
parts = (np.uint32(0) | np.uint32(3 << 4) | (rng.integers(0, 256, n).astype(np.uint32) << 8))
pl = pj.solve(pj.POMCPSolver(tree_queries=100), pomdp, belief_mode=abi.BELIEF_EXTERNAL)
pl.set_belief_particles(parts)
L = lib()
L.pomcp_debug_pf_probe.argtypes = [C.c_void_p, C.c_void_p, C.c_int32]
for it in range(5):
    pl.flush_l2()
    pl.pf_update_sir(5, it & 1, seed=it, n_out=n)
    ms = pl.phase_ms(abi.PHASE_PF)[0]
    buf = np.zeros(1024 * 4, dtype=np.uint64)
    assert L.pomcp_debug_pf_probe(pl._h, buf.ctypes.data_as(C.c_void_p), buf.size) == 0
    t = buf.reshape(-1, 4).astype(np.int64)
    t = t[t[:, 0] > 0]
    t0 = t[:, 0].min()
    rel = (t - t0) / 1e3
    print(f"it {it}: event-to-event {ms * 1e3:.1f} us, {len(t)} CTAs | start {rel[:, 0].min():.1f}..{rel[:, 0].max():.1f} | "
          f"barrier arrival {rel[:, 1].min():.1f}..{rel[:, 1].max():.1f} (phase A per CTA {np.median(rel[:, 1] - rel[:, 0]):.1f}) | "
          f"barrier exit {rel[:, 2].min():.1f}..{rel[:, 2].max():.1f} | end {rel[:, 3].min():.1f}..{rel[:, 3].max():.1f} "
          f"(emit per CTA {np.median(rel[:, 3] - rel[:, 2]):.1f}) us")
pl.close()

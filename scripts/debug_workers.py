import os, sys, time, ctypes as C
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
name, Q, imax, R = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), int(sys.argv[4])
mode = sys.argv[5] if len(sys.argv) > 5 else "batched"
cc = float(sys.argv[6]) if len(sys.argv) > 6 else None
pomdp, c = {"tiger": (pj.TigerPOMDP(), 1000.0), "rs78": (pj.RockSamplePOMDP(7, 8, seed=7008), 20.0), "baby": (pj.BabyPOMDP(), 10.0)}[name]
if cc is not None: c = cc
print("create", flush=True)
solver = pj.POMCPSolver(c=c, rng=1, tree_queries=Q, search_mode=mode, inflight_max=imax, rollouts_per_leaf=R)
gp = pj.solve(solver, pomdp)
print("belief", flush=True)
gp.set_belief_initial()
t = time.time()
print("search", flush=True)
try:
    a, N, V = gp.search(Q)
    print(name, Q, imax, R, "ok", a, N, "%.1f ms" % ((time.time() - t) * 1e3), gp.counters(), flush=True)
except Exception as e:
    print(name, Q, imax, R, "ERR", repr(e), "%.1f ms" % ((time.time() - t) * 1e3), flush=True)
gp.close()

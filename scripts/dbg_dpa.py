import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
which = sys.argv[1]
if which == "baby":
    r = pj.test_solver(pj.POMCPDPWSolver(tree_queries=100), pj.BabyPOMDP(-5.0, -10.0))
    print("baby ok", r)
else:
    pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
    Q = int(sys.argv[2])
    solver = pj.POMCPSolver(c=20.0, rng=3, tree_queries=Q, dpw_action=(10.0, 0.5))
    gp = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    gp.set_belief_initial()
    a, N, V = gp.search(Q)
    print("rs ok", a, N, V, gp.counters())

"""One warmed-up decision (search 1e6 + SIR 2^20) on RockSample(7,8): the ncu target."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
import bench

Q = int(sys.argv[1]) if len(sys.argv) > 1 else bench.TREE_QUERIES
R = int(sys.argv[2]) if len(sys.argv) > 2 else 0  # rollouts per leaf (0 = library default)
pomdp = pj.RockSamplePOMDP(bench.RS_N, bench.RS_K, seed=bench.RS_SEED)
solver = pj.POMCPSolver(c=bench.UCB_C, rng=bench.SOLVER_SEED, tree_queries=Q, search_mode="batched", max_nodes=3 * Q,
                        rollouts_per_leaf=R)
policy = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
parts, _ = bench.initial_particles(pomdp, bench.N_PARTICLES)
policy.set_belief_particles(parts)
for k in range(2):
    policy.reset_counters()
    a, N, V = policy.search(Q)
    cs = policy.counters()  # of this decision: ncu profiles the second one (-s 1)
    policy.pf_update_sir(5, k & 1, seed=k, n_out=bench.N_PARTICLES)
print("ok", a, cs["rollout_steps"], cs["waves"], cs["kernel_launches"])
policy.close()

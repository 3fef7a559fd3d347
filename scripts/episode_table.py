"""Developer probe (GPU box): discounted episode reward on RockSample(7,8), SIR belief of 2^16 particles —
the batched GPU planner against the sequential oracle at equal simulations per step, and at roughly equal
wall-clock per step (north_star criterion 3; the table of DESIGN.md §5)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi
from oracle import oracle as O
from test_gpu_parity import _run_episodes

pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
EPISODES, STEPS, NPART = int(sys.argv[1]) if len(sys.argv) > 1 else 24, 40, 1 << 16


def gpu(Q):
    return lambda ep: pj.solve(pj.POMCPSolver(c=20.0, rng=300 + ep, tree_queries=Q, search_mode="batched"), pomdp,
                               belief_mode=abi.BELIEF_EXTERNAL)


def cpu(Q):
    return lambda ep: O.OraclePlanner(pomdp, O.default_params(c=20.0, seed=300 + ep, tree_queries=Q,
                                                              belief_mode=abi.BELIEF_EXTERNAL))


for name, mk, use_gpu in (("GPU batched 1e4", gpu(10000), True), ("GPU batched 1e5", gpu(100000), True),
                          ("GPU batched 1e6", gpu(1000000), True), ("oracle 1e4", cpu(10000), False),
                          ("oracle 1e5", cpu(100000), False)):
    t = time.time()
    r = _run_episodes(O, pomdp, mk, EPISODES, STEPS, 4242, use_gpu, NPART)
    dt = time.time() - t
    print(f"{name:18s} episodes {len(r):3d}  mean discounted reward {r.mean():6.2f} +- {r.std(ddof=1) / np.sqrt(len(r)):.2f} (SE)  "
          f"wall {dt:6.1f} s", flush=True)

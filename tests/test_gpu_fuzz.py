"""Random solver configurations against the invariants every search must keep (GPU): a seeded sweep over the knobs
that shape the launch - trees per launch, in-flight cap and floor, ramp, rollouts per leaf, budget - on three problems.
Every tree is a complete search: Root.N = Q (src/solver.jl:47-54 runs tree_queries simulations), sum_a N(a) = Q - 1
(the first visit expands the root, src/solver.jl:87-103), the counters add up, values are finite.  (The sweep that
preceded this file found a tree without a worker: tests/test_gpu_multitree.py::test_multitree_cap_below_tree_count.)"""
import numpy as np
import pytest

import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

pytestmark = pytest.mark.gpu


def _problems():
    return [("tiger", pj.TigerPOMDP(), 100.0), ("baby", pj.BabyPOMDP(-5.0, -10.0), 10.0),
            ("rs44", pj.RockSamplePOMDP(4, 4, seed=44), 10.0), ("lightdark", pj.LightDark1D(), 1.0)]


@pytest.mark.parametrize("block", range(4))
def test_random_configurations_keep_the_search_invariants(gpu_lib, block):
    rng = np.random.default_rng(1000 + block)
    problems = _problems()
    for case in range(12):
        name, pomdp, c = problems[int(rng.integers(0, len(problems)))]
        T = int(rng.choice([1, 1, 2, 3, 5, 8]))
        Q = int(rng.choice([1, 2, 3, 17, 257, 1000, 4099, 20000]))
        kw = dict(inflight_max=int(rng.choice([0, 0, 1, 2, 3, 7, 64, 600])),
                  inflight_min=int(rng.choice([0, 0, 1, 5, 200])),
                  inflight_ramp_div=int(rng.choice([0, 0, 1, 8, 1000])),
                  rollouts_per_leaf=int(rng.choice([0, 0, 1, 7, 32, 33, 64, 100, 128])))
        solver = pj.POMCPSolver(c=c, rng=int(rng.integers(1, 1 << 30)), tree_queries=Q, search_mode="batched",
                                n_trees=T, **kw)
        gp = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
        for rep in range(2):  # a second search on a fresh root re-uses the pool and the worker tickets
            gp.set_belief_initial()
            gp.reset_counters()
            a, N, V = gp.search(Q)
            cs = gp.counters()
            ctx = (name, T, Q, kw, rep, cs)
            assert cs["simulations"] + cs["terminal_skips"] == T * Q, ctx
            assert cs["terminal_skips"] == 0, ctx  # initial beliefs hold no terminal state
            assert 0 <= a < len(N) and N.sum() == T * (Q - 1), ctx
            assert np.isfinite(V).all(), ctx
            if T > 1:
                for t in range(T):
                    a_t, rn, n_t, v_t = gp.tree_result(t)
                    assert rn == Q and n_t.sum() == Q - 1, ctx + (t,)
            else:
                assert gp.root_stats()[0] == Q, ctx
        gp.close()

"""GPU tests of root parallelism inside one device (n_trees > 1): several independent trees searched side by
side in ONE launch of the worker kernel, merged like the multi-GPU path merges ranks (src/solver.jl:60-70 is
the merge point: V_a = sum N_a V_a / sum N_a over the trees, then the reference's arg-max).  This is how
BASELINE configs[4] (8 trees x 10^7 simulations) runs on 1/2/4/8 GPUs as 8/4/2/1 trees per GPU."""
import numpy as np
import pytest

import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

pytestmark = pytest.mark.gpu


def _argmax_last(v, exists=None):
    best = -1
    for i in range(len(v)):
        if exists is not None and not exists[i]:
            continue
        if best < 0 or v[i] >= v[best]:
            best = i
    return best


@pytest.mark.parametrize("T,Q", [(2, 40000), (4, 25000), (8, 20000)])
def test_multitree_merge_equals_host_merge(gpu_lib, T, Q):
    """Every tree is a complete search of its own (Root.N = Q, sum_a N(a) = Q - 1, distinct Philox streams);
    the decision the library returns is the host-side merge of the per-tree statistics."""
    pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
    solver = pj.POMCPSolver(c=20.0, rng=11, tree_queries=Q, search_mode="batched", n_trees=T)
    gp = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    rng = np.random.default_rng(0)
    parts = (np.uint32(0) | np.uint32(3 << 4) | (rng.integers(0, 256, 1 << 14).astype(np.uint32) << 8))
    for rep in range(2):  # the second search starts from a fresh root again (a new particle belief)
        gp.set_belief_particles(parts)
        a, N, V = gp.search(Q)
        own = [gp.tree_result(t) for t in range(T)]
        for (_, rn, n_t, v_t) in own:
            assert rn == Q and n_t.sum() == Q - 1 and np.isfinite(v_t).all()
        assert len({tuple(o[2]) for o in own}) == T  # the trees differ
        n = sum(o[2] for o in own)
        nv = sum(o[2] * o[3] for o in own)
        assert np.array_equal(N, n)
        assert np.allclose(V, nv / np.maximum(n, 1), rtol=0, atol=1e-9)
        assert a == _argmax_last(nv / np.maximum(n, 1))
        cs = gp.counters()
        assert cs["simulations"] == (rep + 1) * T * Q
    gp.close()


def test_multitree_matches_oracle_rootparallel(oracle, gpu_lib):
    """4 trees x 50 000 queries against the CPU oracle's root-parallel search with the same streams layout (tree t
    of rank 0 = stream t): both are randomised, so the comparison is statistical - the merged action equals
    the oracle's, the merged best-arm value differs by less than 3 SE + 3 % of the reward scale over 6 seeds."""
    pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
    T, Q = 4, 50000
    vg, vo, same = [], [], 0
    for seed in range(6):
        solver = pj.POMCPSolver(c=20.0, rng=40 + seed, tree_queries=Q, search_mode="batched", n_trees=T)
        gp = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
        gp.set_belief_initial()
        a, N, V = gp.search(Q)
        gp.close()
        sp = oracle.default_params(c=20.0, seed=40 + seed, tree_queries=Q, belief_mode=abi.BELIEF_EXTERNAL)
        rc, a_o, N_o, V_o, _ = oracle.search_rootparallel(pomdp, sp, None, T, T, Q)
        assert rc == 0 and N.sum() == N_o.sum() == T * (Q - 1)
        same += a == a_o
        vg.append(V[a_o])
        vo.append(V_o[a_o])
    vg, vo = np.array(vg), np.array(vo)
    se = np.sqrt(vg.var(ddof=1) / len(vg) + vo.var(ddof=1) / len(vo))
    assert same >= 5, (same, vg, vo)
    assert abs(vg.mean() - vo.mean()) <= 3 * se + 0.3, (vg, vo)


def test_multitree_rejects_what_it_cannot_do(gpu_lib):
    pomdp = pj.TigerPOMDP()
    with pytest.raises(ValueError):
        pj.solve(pj.POMCPSolver(tree_queries=100, search_mode="exact", n_trees=2), pomdp)
    with pytest.raises(ValueError):
        pj.solve(pj.POMCPSolver(tree_queries=100, n_trees=1000), pomdp)
    gp = pj.solve(pj.POMCPSolver(c=100.0, tree_queries=2000, n_trees=2, rng=1), pomdp)
    gp.set_belief_initial()
    a, N, V = gp.search(2000)
    assert N.sum() == 2 * 1999
    with pytest.raises(NotImplementedError):  # several trees share one external belief: no re-rooting
        gp.update_unweighted(a, 0)
    gp.close()


def test_c5_eight_trees_one_gpu_reduced(gpu_lib):
    """BASELINE configs[4] shape on one GPU at a tenth of the size (8 trees x 10^6 on RockSample(11,11)); the full
    8 x 10^7 is bench.py --config c5."""
    pomdp = pj.RockSamplePOMDP(11, 11, seed=11011)
    T, Q = 8, 1_000_000
    solver = pj.POMCPSolver(c=20.0, rng=8, tree_queries=Q, search_mode="batched", n_trees=T, max_nodes=T * 2 * Q + 65536)
    gp = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    gp.set_belief_initial()
    a, N, V = gp.search(Q)
    assert N.sum() == T * (Q - 1) and np.isfinite(V).all() and (N > 0).all()
    for t in range(T):
        a_t, rn, n_t, v_t = gp.tree_result(t)
        assert rn == Q and n_t.sum() == Q - 1
    ms, _ = gp.phase_ms(abi.PHASE_SEARCH)
    cs = gp.counters()
    print(f"C5/10: 8 trees x 1e6: {ms:.1f} ms, {T * Q / ms * 1e3:.3e} sims/s, {cs['rollout_steps'] / ms * 1e3:.3e} rollout "
          f"steps/s, {cs['nodes']} nodes")
    gp.close()


def test_multitree_edge_cases(gpu_lib):
    """Fewer queries than workers, the largest tree count, continuous observations."""
    baby = pj.BabyPOMDP(-5.0, -10.0)
    for T, Q in ((128, 1), (128, 3), (7, 50), (3, 1000)):
        gp = pj.solve(pj.POMCPSolver(c=10.0, rng=1, tree_queries=Q, n_trees=T), baby, belief_mode=abi.BELIEF_EXTERNAL)
        gp.set_belief_initial()
        a, N, V = gp.search(Q)
        assert N.sum() == T * (Q - 1), (T, Q, N)
        for t in (0, T - 1):
            assert gp.tree_result(t)[1] == Q
        gp.close()
    ld = pj.LightDark1D()
    gp = pj.solve(pj.POMCPSolver(c=10.0, rng=2, tree_queries=20000, n_trees=4), ld, belief_mode=abi.BELIEF_EXTERNAL)
    rng = np.random.default_rng(3)
    parts = np.zeros(5000, dtype=ld.state_dtype)
    parts["y"] = rng.normal(2.0, 3.0, 5000)
    gp.set_belief_particles(parts)
    a, N, V = gp.search(20000)
    assert N.sum() == 4 * 19999 and np.isfinite(V).all()
    gp.close()


@pytest.mark.parametrize("kw", [dict(inflight_max=1), dict(inflight_max=3), dict(inflight_max=1, rollouts_per_leaf=32),
                                dict(inflight_max=1, inflight_ramp_div=1000)])
def test_multitree_cap_below_tree_count(gpu_lib, kw):
    """An in-flight cap smaller than n_trees is raised to n_trees: every tree keeps a worker of its own (a tree
    without one was never searched and its empty root went into the merge)."""
    T, Q = 5, 2000
    for pomdp in (pj.RockSamplePOMDP(4, 4, seed=44), pj.BabyPOMDP(-5.0, -10.0)):
        gp = pj.solve(pj.POMCPSolver(c=10.0, rng=5, tree_queries=Q, n_trees=T, **kw), pomdp,
                      belief_mode=abi.BELIEF_EXTERNAL)
        gp.set_belief_initial()
        a, N, V = gp.search(Q)
        cs = gp.counters()
        assert cs["simulations"] + cs["terminal_skips"] == T * Q, (cs, kw)
        for t in range(T):
            assert gp.tree_result(t)[1] == Q, (t, kw)
        assert N.sum() == T * (Q - 1)
        gp.close()

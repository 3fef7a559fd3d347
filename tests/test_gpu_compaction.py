"""GPU tests of the pool compaction that follows a re-root (csrc/reroot.cuh): with
POMCP_COMPACT_ALWAYS=1 every pomcp_update_unweighted rebuilds the node pool and the particle log
around the new root; the search results must stay bit-identical to the (never compacting) oracle."""
import numpy as np
import pytest

import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

from helpers import make_pair, problems

pytestmark = pytest.mark.gpu
PROBS = problems()


def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("name", ["tiger", "baby", "rs44", "rs78"])
def test_compaction_keeps_exact_parity(oracle, gpu_lib, name, monkeypatch):
    monkeypatch.setenv("POMCP_COMPACT_ALWAYS", "1")
    pomdp, c = PROBS[name]
    gp, op = make_pair(oracle, pomdp, c, seed=77, mode="exact", tree_queries=2000)
    gp.set_belief_initial()
    op.set_belief_initial()
    for step in range(6):
        rc, a_o, N_o, V_o = op.search(2000)
        if rc == abi.ERR_ALL_SAMPLES_TERMINAL:  # the episode left the grid: same outcome on the device
            with pytest.raises(pj.AllSamplesTerminal):
                gp.search(2000)
            break
        a_g, N_g, V_g = gp.search(2000)
        assert rc == 0 and a_g == a_o, step
        assert np.array_equal(N_g, N_o) and np.array_equal(_bits(V_g), _bits(V_o)), step
        best_o, best_n = None, 0
        for o in range(pomdp.n_observations):
            n = op.node_stats([(a_o, o)])[0]
            if n > best_n:
                best_o, best_n = o, n
        # two levels below the root must survive the compaction with identical statistics
        sub = [(a_o, best_o)]
        for a2 in range(pomdp.n_actions):
            for o2 in range(pomdp.n_observations):
                ng, _, Ng, Vg = gp.node_stats(sub + [(a2, o2)])
                no, _, No, Vo = op.node_stats(sub + [(a2, o2)])
                assert ng == no
        nodes_before = gp.counters()["nodes"]
        gp.update_unweighted(a_g, best_o)
        assert op.update_unweighted(a_o, best_o) == 0
        cg = gp.counters()
        assert cg["nodes"] <= nodes_before  # the dead part of the pool was reclaimed
        assert np.array_equal(gp.belief_particles(), op.belief_particles())
        rn_g, N1_g, V1_g = gp.root_stats()
        rn_o, N1_o, V1_o = op.root_stats()
        assert rn_g == rn_o and np.array_equal(N1_g, N1_o) and np.array_equal(_bits(V1_g), _bits(V1_o))
        for a2 in range(pomdp.n_actions):
            for o2 in range(pomdp.n_observations):
                ng, pg, Ng, Vg = gp.node_stats([(a2, o2)])
                no, po, No, Vo = op.node_stats([(a2, o2)])
                assert (ng, pg) == (no, po)
                if ng >= 0:
                    assert np.array_equal(Ng, No) and np.array_equal(_bits(Vg), _bits(Vo))
        # structural invariants of the compacted pool
        node_N, act_N, act_V, child = gp.dump_tree()
        assert node_N[0] == rn_g
        for n in range(len(node_N)):
            for a_ in range(pomdp.n_actions):
                kids = child[n, a_][child[n, a_] >= 0]
                assert act_N[n, a_] == node_N[kids].sum()
    gp.close()


def test_compaction_continuous_observations(oracle, gpu_lib, monkeypatch):
    """LightDark (hashed children): replayed uniforms make the first real observation known, so the
    re-root + compaction + re-hash can be followed by a second replayed search and compared."""
    monkeypatch.setenv("POMCP_COMPACT_ALWAYS", "1")
    pomdp = pj.LightDark1D()
    gp, op = make_pair(oracle, pomdp, 1.0, seed=0, mode="exact")
    rng = np.random.default_rng(4)
    Q = 300
    uni = rng.random(Q * 8)
    rets = rng.normal(0.0, 3.0, Q)
    gp.set_belief_initial()
    op.set_belief_initial()
    a_g, N_g, V_g, uu, ru = gp.search_replay(Q, uni, rets)
    rc, a_o, N_o, V_o, uu_o, ru_o = op.search_replay(Q, uni, rets)
    assert rc == 0 and np.array_equal(N_g, N_o) and (uu, ru) == (uu_o, ru_o)
    # query 1 (after the root-expansion query 0): root draw = uni[2:4], tree draws = uni[4:8];
    # at h.N == 1 every criterion is V = 0, so the last action (index 2, move +1) is taken
    y0 = pomdp.init_mean + oracle.lib().orc_normal_from_uniforms(uni[2], uni[3]) * pomdp.init_std
    s = np.zeros(1, dtype=pomdp.state_dtype)
    s["y"] = y0
    sp, obits, r = oracle.model_step(pomdp, s[0], 2, ut=(uni[4], uni[5]), uo=(uni[6], uni[7]))
    o = pj.obs_from_bits(pomdp, obits)
    assert op.node_stats([(2, o)])[0] >= 1 and gp.node_stats([(2, o)])[0] >= 1
    gp.update_unweighted(2, o)
    assert op.update_unweighted(2, o) == 0
    assert gp.belief_particles().tobytes() == op.belief_particles().tobytes()
    uni2 = rng.random(Q * 8)
    a_g, N_g, V_g, uu, ru = gp.search_replay(Q, uni2, rets)
    rc, a_o, N_o, V_o, uu_o, ru_o = op.search_replay(Q, uni2, rets)
    assert rc == 0 and a_g == a_o and np.array_equal(N_g, N_o) and np.array_equal(_bits(V_g), _bits(V_o))
    gp.close()


def test_long_episode_does_not_exhaust_the_pool(oracle, gpu_lib):
    """Without compaction a 60-step tree-reuse episode needs 60 * tree_queries nodes.  Baby keeps about
    90 % of the tree at every re-root, so the live sub-tree converges to roughly 10x tree_queries; the
    pool is sized 16x here and the automatic compaction (at 25 % usage) has to keep it inside."""
    pomdp = pj.BabyPOMDP(-5.0, -10.0)
    Q = 2000
    pl = pj.solve(pj.POMCPSolver(c=10.0, rng=5, tree_queries=Q, search_mode="batched", max_nodes=16 * Q,
                                 max_log=600 * Q), pomdp)
    pl.set_belief_initial()
    env = np.random.default_rng(0)
    s = np.uint32(0)
    for t in range(60):
        a = pl.search(-1)[0]
        s, o, r = oracle.model_step(pomdp, s, a, ut=(env.random(), 0), uo=(env.random(), 0))
        pl.update_unweighted(a, int(o))
        assert pl.counters()["nodes"] <= 16 * Q
    pl.close()


def test_batched_compaction_keeps_the_subtree(gpu_lib, monkeypatch):
    """The batched workers insert with node ids allocated ahead of time, so a child can be older than its
    parent; the compaction walks parent links instead of relying on id order.  After a compacting re-root the
    kept sub-tree must be exactly the old one: same root visit count, same root statistics, N(ha) = sum_o N(hao)
    everywhere, and no id lost to insertion races (ids in use <= nodes + one spare per worker)."""
    monkeypatch.setenv("POMCP_COMPACT_ALWAYS", "1")
    pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
    Q = 200_000
    pl = pj.solve(pj.POMCPSolver(c=20.0, rng=9, tree_queries=Q, search_mode="batched", max_nodes=4 * Q,
                                 max_log=40 * Q), pomdp)
    pl.set_belief_initial()
    a, N, V = pl.search(Q)
    cs = pl.counters()
    assert cs["nodes"] <= cs["simulations"] + 2368 + 1
    node_N, act_N, act_V, child = pl.dump_tree()
    kid_N = np.where(child >= 0, node_N[np.maximum(child, 0)], 0).sum(axis=2)
    assert np.array_equal(kid_N, act_N)
    # the child ids are NOT monotone along paths any more (that is the point of the test)
    parent_of = np.full(node_N.shape[0], -1, dtype=np.int64)
    hs, as_, os_ = np.nonzero(child >= 0)
    parent_of[child[hs, as_, os_]] = hs
    linked = np.nonzero(parent_of >= 0)[0]
    assert (parent_of[linked] > linked).any()
    # re-root at the most visited child of the chosen action
    o = int(np.argmax([pl.node_stats([(a, k)])[0] for k in range(pomdp.n_observations)]))
    n_before, _, N_before, V_before = pl.node_stats([(a, o)])
    sub_before = {(a2, o2): pl.node_stats([(a, o), (a2, o2)])[0] for a2 in range(pomdp.n_actions) for o2 in range(3)}
    pl.update_unweighted(a, o)
    rn, N1, V1 = pl.root_stats()
    assert rn == n_before and np.array_equal(N1, N_before) and np.array_equal(_bits(V1), _bits(V_before))
    for (a2, o2), n in sub_before.items():
        assert pl.node_stats([(a2, o2)])[0] == n
    node_N, act_N, act_V, child = pl.dump_tree()
    kid_N = np.where(child >= 0, node_N[np.maximum(child, 0)], 0).sum(axis=2)
    assert np.array_equal(kid_N, act_N)
    assert pl.counters()["nodes"] < cs["nodes"]
    a2, N2, V2 = pl.search(Q)  # and the tree keeps growing from there
    assert pl.root_stats()[0] == rn + Q
    pl.close()


@pytest.mark.parametrize("name", ["tiger", "rs44"])
def test_compaction_with_action_widening_keeps_exact_parity(oracle, gpu_lib, name, monkeypatch):
    """Action progressive widening keeps a mask of the existing ActNodes per belief node; the compaction moves it with
    the node (round 2: before, a planner with action widening never compacted).  Exact-mode episodes with a compaction
    after every re-root stay bit-identical to the never-compacting oracle."""
    monkeypatch.setenv("POMCP_COMPACT_ALWAYS", "1")
    pomdp, c = PROBS[name]
    gp, op = make_pair(oracle, pomdp, c, seed=5, mode="exact", tree_queries=1500, dpw_action=(2.0, 0.3))
    gp.set_belief_initial()
    op.set_belief_initial()
    for step in range(5):
        rc, a_o, N_o, V_o = op.search(1500)
        if rc == abi.ERR_ALL_SAMPLES_TERMINAL:
            with pytest.raises(pj.AllSamplesTerminal):
                gp.search(1500)
            break
        a_g, N_g, V_g = gp.search(1500)
        assert rc == 0 and a_g == a_o and np.array_equal(N_g, N_o) and np.array_equal(_bits(V_g), _bits(V_o)), step
        best_o, best_n = None, 0
        for o in range(pomdp.n_observations):
            n = op.node_stats([(a_o, o)])[0]
            if n > best_n:
                best_o, best_n = o, n
        if best_o is None:
            break
        nodes_before = gp.counters()["nodes"]
        gp.update_unweighted(a_g, best_o)
        assert op.update_unweighted(a_o, best_o) == 0
        assert gp.counters()["nodes"] < nodes_before  # the pool was compacted
        assert np.array_equal(gp.belief_particles(), op.belief_particles())
        rn_g, N1_g, V1_g = gp.root_stats()
        rn_o, N1_o, V1_o = op.root_stats()
        assert rn_g == rn_o and np.array_equal(N1_g, N1_o) and np.array_equal(_bits(V1_g), _bits(V1_o))
    gp.close()

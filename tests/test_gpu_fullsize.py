"""GPU tests at the sizes BASELINE.json names (configs C1..C5 of SURVEY.md §8): where the oracle
finishes in seconds the comparison is direct; at 10^6 / 10^7 queries it goes through
size-independent properties of the reference algorithm:
Root.N = tree_queries, sum_a N(a) = tree_queries - 1 (the root expansion consumes one query),
N(ha) = sum_o N(hao), particles pushed at a node = its visit count, SIR indices bit-exact."""
import numpy as np
import pytest

import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

from helpers import make_pair

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


def _check_counts(node_N, act_N, child):
    """Vectorised Q10: N(ha) = sum_o N(hao) for every action node of the dump."""
    kid_N = np.where(child >= 0, node_N[np.maximum(child, 0)], 0).sum(axis=2)
    bad = np.nonzero(kid_N != act_N)
    assert bad[0].size == 0, (bad[0][:5], bad[1][:5])


# ----------------------------------------------------------------------------- C1
def test_c1_tiger_1000_queries_unweighted_episode(oracle, gpu_lib):
    """C1: TigerPOMDP, tree_queries = 1000, random rollouts, unweighted particle updater, a seeded
    20-step episode.  Exact mode: every decision, every root statistic and every re-rooted particle
    set equals the sequential oracle's, bit for bit, over 5 seeds."""
    pomdp = pj.TigerPOMDP()
    for seed in range(5):
        gp, op = make_pair(oracle, pomdp, 100.0, seed=seed, mode="exact", tree_queries=1000)
        gp.set_belief_initial()
        op.set_belief_initial()
        rng = np.random.default_rng(seed)
        s = int(rng.random() < 0.5)
        for t in range(20):
            a_g, N_g, V_g = gp.search(1000)
            rc, a_o, N_o, V_o = op.search(1000)
            assert rc == 0 and a_g == a_o and np.array_equal(N_g, N_o) and np.array_equal(_bits(V_g), _bits(V_o))
            sp, o, r = oracle.model_step(pomdp, s, a_g, ut=(rng.random(), 0.0), uo=(rng.random(), 0.0))
            rc = op.update_unweighted(a_g, int(o))
            if rc == abi.ERR_PARTICLE_DEPLETION:
                with pytest.raises(pj.ParticleDepletion):
                    gp.update_unweighted(a_g, int(o))
                break
            gp.update_unweighted(a_g, int(o))
            assert gp.belief_particles().tobytes() == op.belief_particles().tobytes()
            s = int(sp)
        gp.close()


# ----------------------------------------------------------------------------- C2
def test_c2_baby_1e5_queries_1e4_particles_sir(oracle, gpu_lib):
    """C2: BabyPOMDP, tree_queries = 10^5, 10^4 particles, SIR update.  Search: structural invariants
    at full size; value of the chosen action not worse than the sequential oracle's on the same
    belief.  (A single POMCP run with c = 10 can lock onto either action at P(hungry) = 0.1 - the arm
    it neglects keeps its first noisy rollout estimates - so the comparison is one-sided and between
    means over 4 seeds: mean max_a V(GPU) >= mean max_a V(oracle) - 0.5, i.e. 3% of the reward
    scale, and inside the band [-21, -15] that contains the optimal value.)  SIR: the resampled
    particle set is bit-identical to the oracle's."""
    pomdp = pj.BabyPOMDP(-5.0, -10.0)
    Q, n = 100000, 10000
    rng = np.random.default_rng(2)
    parts = (rng.random(n) < 0.1).astype(np.uint32)
    best_g, best_o = [], []
    for seed in range(4):
        gp, op = make_pair(oracle, pomdp, 10.0, seed=17 + seed, mode="batched", belief_mode=abi.BELIEF_EXTERNAL,
                           tree_queries=Q)
        gp.set_belief_particles(parts)
        op.set_belief_particles(parts)
        a_g, N_g, V_g = gp.search(Q)
        rc, a_o, N_o, V_o = op.search(Q)
        assert rc == 0
        assert gp.root_stats()[0] == Q and N_g.sum() == Q - 1
        if seed == 0:
            node_N, act_N, act_V, child = gp.dump_tree()
            _check_counts(node_N, act_N, child)
        best_g.append(V_g[a_g])
        best_o.append(V_o[a_o])
        if seed < 3:
            gp.close()
    assert np.mean(best_g) >= np.mean(best_o) - 0.5, (best_g, best_o)
    assert -21.0 < np.mean(best_g) < -15.0, best_g
    for k, (a, o) in enumerate([(0, 0), (0, 1), (1, 0), (0, 1)]):  # alternating (a, o) script
        gp.pf_update_sir(a, o, seed=100 + k, n_out=n)
        assert op.pf_update_sir(a, o, seed=100 + k, n_out=n, mode=0) == 0
        assert gp.belief_particles().tobytes() == op.belief_particles().tobytes()
    gp.close()


# ----------------------------------------------------------------------------- C3
def test_c3_rocksample_7_8_1e6_queries(oracle, gpu_lib):
    """C3: RockSample(7,8), tree_queries = 10^6 per action.  Properties at full size: Root.N = Q,
    sum_a N(a) = Q - 1, N(ha) = sum_o N(hao) over all ~1.7 M nodes, one rollout batch per expanded
    non-root node, and the decision of the sequential oracle (3.2 s of CPU) on the same belief."""
    pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
    Q = 1_000_000
    gp, op = make_pair(oracle, pomdp, 20.0, seed=5, mode="batched", belief_mode=abi.BELIEF_EXTERNAL, tree_queries=Q,
                       max_nodes=3 * Q)
    rng = np.random.default_rng(0)
    parts = (np.uint32(0) | np.uint32(3 << 4) | (rng.integers(0, 256, 1 << 16).astype(np.uint32) << 8))
    gp.set_belief_particles(parts)
    op.set_belief_particles(parts)
    a_g, N_g, V_g = gp.search(Q)
    assert gp.root_stats()[0] == Q and N_g.sum() == Q - 1
    cs = gp.counters()
    assert cs["simulations"] == Q and cs["terminal_skips"] == 0
    assert cs["rollouts"] % 64 == 0 and cs["rollouts"] // 64 <= cs["nodes"] - 1  # R = 64 per leaf (batched default)
    assert cs["rollout_steps"] <= cs["rollouts"] * 89  # steps <= ceil(log(eps)/log(gamma) - depth), depth >= 1
    node_N, act_N, act_V, child = gp.dump_tree()
    _check_counts(node_N, act_N, child)
    assert np.isfinite(act_V).all()
    rc, a_o, N_o, V_o = op.search(Q)
    assert rc == 0 and a_g == a_o, (a_g, a_o, V_g, V_o)
    assert abs(V_g[a_o] - V_o[a_o]) < 0.2 * abs(V_o[a_o]) + 0.5, (V_g, V_o)
    gp.close()


def test_c3_values_vs_oracle(oracle, gpu_lib):
    """north_star criterion 2 at the headline size: RockSample(7,8), 10^6 queries, 6 seeds, the batched planner
    against the sequential oracle on the same seeds (one oracle tree per host thread).  Stated tolerances:
      * the batched planner picks the oracle's modal root action in >= 5 of 6 runs, at every setting;
      * with 64 simulations in flight the mean value of that action differs from the oracle's by no more than
        3 standard errors of the difference + 3 % of the reward scale (10) - the rule of the 10^5-query test;
      * with the defaults (8 simulations in flight per SM = 1184, 64 rollouts per leaf) the batched value is LOWER
        than the oracle's (never higher beyond noise) by no more than 3 SE + 15 % of the reward scale.
    The second number is the measured price of tree parallelism, not noise (scripts/probe.py knob, 12 seeds:
    -0.06 +- 0.16 at 64 in flight, -0.58 +- 0.13 at 256, -1.17 +- 0.17 at 1184; the oracle with 64 rollouts per leaf
    scores 10.17 +- 0.07 against 10.06 +- 0.13 with one, so the estimator is not the cause): a simulation cannot see
    the results of the ones still in flight, which costs the tree some exploitation.  At equal wall time the wide
    setting wins by far (probe.py equal-time: 8.8 against 6.2 at 256 and 3.4 at 64 in flight in 14 ms), the root
    action is the oracle's in 12 of 12 runs at every setting, and the episode reward does not differ
    (tests/test_gpu_episode.py)."""
    from concurrent.futures import ThreadPoolExecutor
    pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
    Q, seeds = 1_000_000, [60 + i for i in range(6)]

    def orc(seed):
        op = oracle.OraclePlanner(pomdp, oracle.default_params(c=20.0, seed=seed, tree_queries=Q))
        op.set_belief_initial()
        rc, a, N, V = op.search(Q)
        op.close()
        assert rc == 0
        return a, V

    def gpu(seed, **kw):
        gp = pj.solve(pj.POMCPSolver(c=20.0, rng=seed, tree_queries=Q, search_mode="batched", max_nodes=3 * Q, **kw), pomdp,
                      belief_mode=abi.BELIEF_EXTERNAL)
        gp.set_belief_initial()
        a, N, V = gp.search(Q)
        assert N.sum() == Q - 1
        gp.close()
        return a, V

    with ThreadPoolExecutor(len(seeds)) as ex:
        fut = [ex.submit(orc, s) for s in seeds]  # CPU: runs while the GPU searches below
        res_wide = [gpu(s) for s in seeds]
        res_narrow = [gpu(s, inflight_max=64) for s in seeds]
        res_o = [f.result() for f in fut]
    modal = int(np.bincount([a for a, _ in res_o]).argmax())
    vo = np.array([v[modal] for _, v in res_o])
    for name, res, slack, one_sided in (("64 in flight", res_narrow, 0.03 * 10.0, False), ("defaults", res_wide, 0.15 * 10.0, True)):
        assert sum(a == modal for a, _ in res) >= 5, (name, [a for a, _ in res], [a for a, _ in res_o])
        vg = np.array([v[modal] for _, v in res])
        se = np.sqrt(vg.var(ddof=1) / len(vg) + vo.var(ddof=1) / len(vo))
        print(f"C3 values at 1e6, {name}: batched {vg.mean():.3f} +- {vg.std(ddof=1) / np.sqrt(len(vg)):.3f}, oracle "
              f"{vo.mean():.3f} +- {vo.std(ddof=1) / np.sqrt(len(vo)):.3f}")
        d = vg.mean() - vo.mean()
        assert d <= 3 * se + (0.03 * 10.0 if one_sided else slack), (name, vg, vo, se)
        assert d >= -(3 * se + slack), (name, vg, vo, se)


# ----------------------------------------------------------------------------- C4
@pytest.mark.parametrize("streaming", [False, True])
def test_c4_lightdark_1e6_particles_sir(oracle, gpu_lib, streaming, monkeypatch):
    """C4: LightDark1D, 10^6 particles y ~ N(2, 3^2), SIR update with a = +1, o = 3.0, n_out = 10^6:
    the resampled set is bit-identical to the oracle's sequential walk (fused and streaming paths);
    then a 10^5-query plain POMCP search on it keeps the counting invariants (continuous
    observations: one new node per simulation, depth 1)."""
    if streaming:
        monkeypatch.setenv("POMCP_PF_STREAMING", "1")
    pomdp = pj.LightDark1D()
    n, Q = 1_000_000, 100_000
    gp, op = make_pair(oracle, pomdp, 1.0, seed=4, mode="batched", belief_mode=abi.BELIEF_EXTERNAL, tree_queries=Q)
    rng = np.random.default_rng(11)
    parts = np.zeros(n, dtype=pomdp.state_dtype)
    parts["y"] = rng.normal(2.0, 3.0, n)
    gp.set_belief_particles(parts)
    op.set_belief_particles(parts)
    gp.pf_update_sir(2, 3.0, seed=77, n_out=n)
    assert op.pf_update_sir(2, 3.0, seed=77, n_out=n, mode=0) == 0
    got, want = gp.belief_particles(), op.belief_particles()
    assert got.tobytes() == want.tobytes()
    # independent check of the posterior: importance-weighted mean of the propagated prior (numpy)
    yp = parts["y"] + 1.0
    sd = np.abs(yp - 5.0) / np.sqrt(2.0) + 1e-2
    w = np.exp(-(3.0 - yp) ** 2 / (2.0 * sd * sd)) / (sd * np.sqrt(2.0 * np.pi))
    assert abs(got["y"].mean() - float((w * yp).sum() / w.sum())) < 0.02
    if not streaming:
        a_g, N_g, V_g = gp.search(Q)
        assert gp.root_stats()[0] == Q and N_g.sum() == Q - 1
        cs = gp.counters()
        assert cs["max_depth_seen"] == 1 and cs["nodes"] == Q
    gp.close()


def test_sir_rocksample_2p26_particles_tiled_path(oracle, gpu_lib):
    """The size bench.py quotes the SIR roofline on: 2^26 RockSample(7,8) particles through the tiled path (28 tiles
    per CTA).  The resampled set is bit-identical to the oracle's sequential walk, for a sensing action (copies are the
    input states) and for a move (copies are propagated), with a block of terminal particles in front."""
    pomdp = pj.RockSamplePOMDP(7, 8, seed=7008)
    n = 1 << 26
    gp, op = make_pair(oracle, pomdp, 20.0, seed=4, mode="batched", belief_mode=abi.BELIEF_EXTERNAL, tree_queries=100)
    rng = np.random.default_rng(26)
    parts = (rng.integers(0, 7, n).astype(np.uint32) | (rng.integers(0, 7, n).astype(np.uint32) << 4)
             | (rng.integers(0, 256, n).astype(np.uint32) << 8))
    parts[:5000] |= np.uint32(abi.RS_TERMINAL)
    for k, (a, o, n_out) in enumerate([(5, 0, n), (1, 2, n // 2 + 3)]):
        gp.set_belief_particles(parts)
        op.set_belief_particles(parts)
        gp.pf_update_sir(a, o, seed=900 + k, n_out=n_out)
        assert gp.phase_ms(abi.PHASE_PF)[1] == 2  # two launches: weigh + chunk sums, emit
        assert op.pf_update_sir(a, o, seed=900 + k, n_out=n_out, mode=0) == 0
        got, want = gp.belief_particles(), op.belief_particles()
        assert len(got) == n_out and got.tobytes() == want.tobytes(), (k, a, o)
    gp.close()


# ----------------------------------------------------------------------------- C5 (one rank's share)
def test_c5_rocksample_11_11_1e7_sims_one_tree(gpu_lib):
    """C5: RockSample(11,11), one tree of 10^7 simulations (the per-GPU share of the 8-tree job; the
    NCCL merge is covered by test_gpu_multirank.py).  Properties: Root.N = Q, sum_a N(a) = Q - 1, the
    pool holds the tree (no POOL_EXHAUSTED), every root action was tried, values finite."""
    pomdp = pj.RockSamplePOMDP(11, 11, seed=11011)
    Q = 10_000_000
    solver = pj.POMCPSolver(c=20.0, rng=8, tree_queries=Q, search_mode="batched", max_nodes=2 * Q + 65536)
    gp = pj.solve(solver, pomdp, belief_mode=abi.BELIEF_EXTERNAL)
    gp.set_belief_initial()
    a, N, V = gp.search(Q)
    assert gp.root_stats()[0] == Q and N.sum() == Q - 1
    assert (N > 0).all() and np.isfinite(V).all() and 0 <= a < 16
    cs = gp.counters()
    assert cs["simulations"] == Q and cs["nodes"] <= 2 * Q + 65536
    ms, _ = gp.phase_ms(abi.PHASE_SEARCH)
    print(f"C5 one tree: {Q / ms * 1e3:.3e} sims/s, {cs['rollout_steps'] / ms * 1e3:.3e} rollout steps/s, "
          f"{cs['nodes']} nodes, depth <= {cs['max_depth_seen']}")
    gp.close()

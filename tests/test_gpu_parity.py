"""GPU parity tests: every CUDA path is called through the C ABI (via the host mirror) and compared
with the CPU oracle on the same seeded inputs.  Integer / index / count results must be bit-exact;
fp64 values that go through identical IEEE operation sequences must be bit-exact too; the batched
search (atomics, waves in flight) is compared statistically with the tolerance written in the test."""
import numpy as np
import pytest

import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

from helpers import make_pair, problems, random_states

pytestmark = pytest.mark.gpu

PROBS = problems()


def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


# ----------------------------------------------------------------------------- K3 rollouts
@pytest.mark.parametrize("name", list(PROBS))
def test_rollout_batch_bit_exact(oracle, gpu_lib, name):
    pomdp, c = PROBS[name]
    gp, _ = make_pair(oracle, pomdp, c, seed=3)
    rng = np.random.default_rng(1)
    states = random_states(pomdp, 20000, rng)
    for steps in (0, 1, 7, 89):
        ret_g, ns_g = gp.rollout_batch(states, steps, seed=0xDEADBEEF12345)
        ret_o, ns_o = oracle.rollout_batch(pomdp, states, steps, seed=0xDEADBEEF12345)
        assert np.array_equal(ns_g, ns_o), (name, steps)
        assert np.array_equal(_bits(ret_g), _bits(ret_o)), (name, steps, np.abs(ret_g - ret_o).max())
    assert ns_g.max() <= 89 and (ns_g > 0).any()
    gp.close()


@pytest.mark.parametrize("n,k", [(11, 11), (10, 12), (13, 14), (15, 16), (2, 1), (1, 0)])
def test_rollout_and_search_large_rocksample_bit_exact(oracle, gpu_lib, n, k):
    """The RockSample rollout table at its edges: 16 actions exactly (k = 11: no clamping of the action column), more
    than 16 actions (k >= 12: actions from 15 on share the sensing column), the largest instance the ABI admits (15 x 15
    grid, 16 rocks: the rock mask fills bits 14..29 of a table entry), and degenerate grids.  Rollout returns, step
    counts and an exact-mode search with 64 rollouts per leaf (two trajectories per lane) equal the oracle's."""
    pomdp = pj.RockSamplePOMDP(n, k, seed=100 * n + k)
    gp, op = make_pair(oracle, pomdp, 20.0, seed=5, mode="exact", tree_queries=300, rollouts_per_leaf=64)
    rng = np.random.default_rng(n + k)
    states = random_states(pomdp, 8000, rng)
    for steps in (1, 8, 9, 90):
        ret_g, ns_g = gp.rollout_batch(states, steps, seed=0xABCDEF + steps)
        ret_o, ns_o = oracle.rollout_batch(pomdp, states, steps, seed=0xABCDEF + steps)
        assert np.array_equal(ns_g, ns_o), (n, k, steps)
        assert np.array_equal(_bits(ret_g), _bits(ret_o)), (n, k, steps, np.abs(ret_g - ret_o).max())
    gp.set_belief_initial()
    op.set_belief_initial()
    a_g, N_g, V_g = gp.search(300)
    rc, a_o, N_o, V_o = op.search(300)
    assert rc == 0 and a_g == a_o and np.array_equal(N_g, N_o) and np.array_equal(_bits(V_g), _bits(V_o))
    gp.close()
    # batched workers on the same instance (one-warp workers, the table-driven tree step): counting invariants
    bp = pj.solve(pj.POMCPSolver(c=20.0, rng=6, tree_queries=20000), pomdp)
    bp.set_belief_initial()
    a, N, V = bp.search(20000)
    assert bp.counters()["simulations"] == 20000 and N.sum() == 20000 - 1 and np.isfinite(V).all()
    bp.close()


# ----------------------------------------------------------------------------- K6 weights
@pytest.mark.parametrize("name", list(PROBS))
def test_pf_weights_bit_exact(oracle, gpu_lib, name):
    pomdp, c = PROBS[name]
    gp, _ = make_pair(oracle, pomdp, c, seed=3)
    rng = np.random.default_rng(2)
    states = random_states(pomdp, 30000, rng)
    obs_list = [3.0, -1.25, 7.5] if pomdp.continuous_obs else list(range(pomdp.n_observations))
    for a in range(pomdp.n_actions):
        for o in obs_list[:2]:
            w_g, sp_g = gp.pf_weights(states, a, o, seed=77)
            w_o, sp_o = oracle.pf_weights(pomdp, states, a, o, seed=77)
            assert np.array_equal(_bits(w_g), _bits(w_o)), (name, a, o)
            assert sp_g.tobytes() == sp_o.tobytes(), (name, a, o)
    gp.close()


# ----------------------------------------------------------------------------- K7/K8 resampling
def _weight_cases(rng):
    yield "uniform", np.ones(1000), 0.37, 1000
    yield "random", rng.random(5000), 0.9123, 5000
    yield "skewed", rng.random(4097) ** 12, 0.5, 3000
    yield "zeros", np.where(rng.random(7000) < 0.8, 0.0, rng.random(7000)), 0.001, 9000
    yield "leading_zero_r0", np.concatenate([[0.0, 0.0], rng.random(100)]), 0.0, 64
    yield "single", np.array([0.3]), 0.99, 17
    yield "one_heavy", np.concatenate([rng.random(3000) * 1e-9, [5.0], rng.random(2000) * 1e-9]), 0.25, 4000
    yield "tiny", rng.random(2500) * 1e-200, 0.6, 2500
    yield "big", rng.random(300000), 0.4242, 300000
    yield "ragged", rng.random(1025), 0.75, 7
    yield "max_single_grid", rng.random(148 * 2 * 4096), 0.123, 1 << 20   # largest set of the fused kernel
    yield "beyond_single_grid", rng.random(1_500_000), 0.321, 1_000_000   # falls back to the streaming kernels
    yield "heavy_runs", np.where(rng.random(20000) < 0.002, 1.0, 1e-12), 0.5, 400000  # few particles, long copy runs


@pytest.mark.parametrize("streaming", [False, True])
def test_resample_fixed_bit_exact(oracle, gpu_lib, streaming, monkeypatch):
    """Fixed-point systematic resampling: fused single-launch kernel and the three streaming kernels
    (POMCP_PF_STREAMING=1 forces the latter) give the oracle's indices bit for bit."""
    if streaming:
        monkeypatch.setenv("POMCP_PF_STREAMING", "1")
    pomdp, c = PROBS["baby"]
    gp, _ = make_pair(oracle, pomdp, c, seed=3)
    rng = np.random.default_rng(5)
    for tag, w, r, n_out in _weight_cases(rng):
        idx_g = gp.resample_systematic(w, r, n_out, mode=0)
        rc, idx_o = oracle.resample_systematic(w, r, n_out, fixed=True)
        assert rc == 0
        assert np.array_equal(idx_g, idx_o), tag
        assert np.all(np.diff(idx_g) >= 0), tag  # sortedness
        # agreement with the upstream fp64 running-sum semantics (not guaranteed bit-exact: re-association)
        rc, idx_f = oracle.resample_systematic(w, r, n_out, fixed=False)
        assert rc == 0
        assert np.mean(idx_f != idx_g) <= 1e-3, (tag, np.mean(idx_f != idx_g))
    gp.close()


def test_resample_sequential_fp64_bit_exact(oracle, gpu_lib):
    """Verification mode: the upstream running fp64 sums on one device thread."""
    pomdp, c = PROBS["baby"]
    gp, _ = make_pair(oracle, pomdp, c, seed=3)
    rng = np.random.default_rng(6)
    for tag, w, r, n_out in _weight_cases(rng):
        if len(w) > 50000:
            continue
        idx_g = gp.resample_systematic(w, r, n_out, mode=1)
        rc, idx_o = oracle.resample_systematic(w, r, n_out, fixed=False)
        assert rc == 0
        assert np.array_equal(idx_g, idx_o), tag
    gp.close()


def test_resample_zero_weight_is_depletion(oracle, gpu_lib):
    gp, _ = make_pair(oracle, PROBS["baby"][0], 10.0, seed=3)
    with pytest.raises(pj.ParticleDepletion):
        gp.resample_systematic(np.zeros(100), 0.5, 100, mode=0)
    rc, _ = oracle.resample_systematic(np.zeros(100), 0.5, 100, fixed=True)
    assert rc == abi.ERR_PARTICLE_DEPLETION
    gp.close()


# ----------------------------------------------------------------------------- exact search
@pytest.mark.parametrize("name,queries,R", [("tiger", 1000, 1), ("baby", 3000, 1), ("rs78", 4000, 1), ("rs44", 4000, 1),
                                            ("lightdark", 1500, 1), ("baby", 2000, 32), ("rs78", 3000, 32),
                                            ("tiger", 1000, 5), ("lightdark", 1000, 32), ("rs78", 2000, 64), ("rs44", 2000, 45),
                                            ("baby", 1500, 64)])
def test_exact_search_bit_exact(oracle, gpu_lib, name, queries, R):
    """Same Philox streams => identical trees: N counts, V bits, structure.  R = rollouts averaged per
    leaf (R = 1 is the reference; lane r runs rollout r and the returns are added in butterfly order)."""
    pomdp, c = PROBS[name]
    gp, op = make_pair(oracle, pomdp, c, seed=11, mode="exact", tree_queries=queries, rollouts_per_leaf=R)
    gp.set_belief_initial()
    op.set_belief_initial()
    a_g, N_g, V_g = gp.search(queries)
    rc, a_o, N_o, V_o = op.search(queries)
    assert rc == 0
    assert np.array_equal(N_g, N_o), (N_g, N_o)
    assert np.array_equal(_bits(V_g), _bits(V_o)), (V_g, V_o)
    assert a_g == a_o
    tg, to = gp.dump_tree(), op.dump_tree()
    assert np.array_equal(tg[0], to[0])
    assert np.array_equal(tg[1], to[1])
    assert np.array_equal(_bits(tg[2]), _bits(to[2]))
    if to[3] is not None:
        assert np.array_equal(tg[3], to[3])
    cg, co = gp.counters(), op.counters()
    for k in ("simulations", "rollouts", "rollout_steps", "tree_steps", "nodes", "log_entries", "max_depth_seen"):
        assert cg[k] == co[k], (k, cg[k], co[k])
    gp.close()


@pytest.mark.parametrize("R,queries", [(1, 3000), (32, 1500)])
def test_fwc_rollout_exact_bit_exact(oracle, gpu_lib, R, queries):
    """RolloutEstimator(FeedWhenCrying()) on Baby (test/runtests.jl:13-19): the rollout policy feeds iff
    the previous observation was crying, starting from the leaf's observation label
    (extract_belief for PreviousObservationUpdater, src/rollout.jl:105-106).  Bit-exact tree."""
    pomdp, c = PROBS["baby"]
    est = pj.RolloutEstimator(pj.FeedWhenCrying())
    gp, op = make_pair(oracle, pomdp, c, seed=23, mode="exact", tree_queries=queries, rollouts_per_leaf=R,
                       estimate_value=est)
    gp.set_belief_initial()
    op.set_belief_initial()
    for Q in (queries, queries // 2):  # second decision continues on the same tree (next epoch)
        a_g, N_g, V_g = gp.search(Q)
        rc, a_o, N_o, V_o = op.search(Q)
        assert rc == 0 and a_g == a_o
        assert np.array_equal(N_g, N_o), (N_g, N_o)
        assert np.array_equal(_bits(V_g), _bits(V_o)), (V_g, V_o)
    tg, to = gp.dump_tree(), op.dump_tree()
    for i in range(2):
        assert np.array_equal(tg[i], to[i])
    assert np.array_equal(_bits(tg[2]), _bits(to[2]))
    cg, co = gp.counters(), op.counters()
    for k in ("simulations", "rollouts", "rollout_steps", "tree_steps", "nodes"):
        assert cg[k] == co[k], (k, cg[k], co[k])
    gp.close()


def test_fwc_rollout_batched_beats_random_rollouts(oracle, gpu_lib):
    """Batched mode with FeedWhenCrying rollouts: leaf estimates come from a near-optimal policy, so
    the root values must be higher than with uniform-random rollouts (notebooks/Sanity_Checks.ipynb:
    FWC alone -15.1 vs random -31.3) and agree with the sequential oracle within 3 SE of its own
    seed-to-seed spread (tolerance 0.6 on values around -17)."""
    pomdp, c = PROBS["baby"]
    est = pj.RolloutEstimator(pj.FeedWhenCrying())
    Q = 40000
    gp, op = make_pair(oracle, pomdp, c, seed=5, mode="batched", tree_queries=Q, estimate_value=est)
    gp.set_belief_initial()
    a_g, N_g, V_g = gp.search(Q)
    assert N_g.sum() == Q - 1
    gp.close()
    gp2, _ = make_pair(oracle, pomdp, c, seed=5, mode="batched", tree_queries=Q)
    gp2.set_belief_initial()
    _, _, V_rand = gp2.search(Q)
    gp2.close()
    sp = op._sp
    sp.search_mode = abi.MODE_EXACT
    sp.rollouts_per_leaf = 1
    op2 = oracle.OraclePlanner(pomdp, sp)
    op2.set_belief_initial()
    rc, a_o, N_o, V_o = op2.search(Q)
    assert rc == 0
    assert a_g == a_o == 0
    assert V_g.max() > V_rand.max()
    assert abs(V_g[a_o] - V_o[a_o]) < 0.6, (V_g, V_o)


def test_fwc_rollout_rejected_outside_baby(gpu_lib):
    with pytest.raises(NotImplementedError):
        pj.solve(pj.POMCPSolver(estimate_value=pj.RolloutEstimator(pj.FeedWhenCrying())), pj.TigerPOMDP())


def test_fo_rollout_and_fo_value_bit_exact(oracle, gpu_lib):
    """FORollout(FeedWhenCrying()) on Baby and FOValue(table) on Baby / RockSample(4,4): exact-mode trees equal
    the oracle's bit for bit; batched mode keeps the counting invariants."""
    baby, c = PROBS["baby"]
    gp, op = make_pair(oracle, baby, c, seed=29, mode="exact", tree_queries=2000, estimate_value=pj.FORollout(pj.FeedWhenCrying()))
    gp.set_belief_initial()
    op.set_belief_initial()
    a_g, N_g, V_g = gp.search(2000)
    rc, a_o, N_o, V_o = op.search(2000)
    assert rc == 0 and a_g == a_o and np.array_equal(N_g, N_o) and np.array_equal(_bits(V_g), _bits(V_o))
    assert gp.counters()["rollout_steps"] == op.counters()["rollout_steps"]
    gp.close()
    rng = np.random.default_rng(9)
    for name, n_states in (("baby", 2), ("rs44", 1 << 12)):
        pomdp, c = PROBS[name]
        table = rng.normal(-5.0, 3.0, n_states)
        gp, op = make_pair(oracle, pomdp, c, seed=31, mode="exact", tree_queries=2500, estimate_value=pj.FOValue(table))
        assert op.set_state_values(table) == 0
        gp.set_belief_initial()
        op.set_belief_initial()
        a_g, N_g, V_g = gp.search(2500)
        rc, a_o, N_o, V_o = op.search(2500)
        assert rc == 0 and a_g == a_o and np.array_equal(N_g, N_o) and np.array_equal(_bits(V_g), _bits(V_o))
        tg, to = gp.dump_tree(), op.dump_tree()
        assert np.array_equal(tg[1], to[1]) and np.array_equal(_bits(tg[2]), _bits(to[2]))
        assert gp.counters()["rollout_steps"] == 0
        gp.close()
        gp, _ = make_pair(oracle, pomdp, c, seed=31, mode="batched", tree_queries=30000, estimate_value=pj.FOValue(table))
        gp.set_belief_initial()
        a, N, V = gp.search(30000)
        assert N.sum() == 29999 and np.isfinite(V).all()
        gp.close()
    with pytest.raises(NotImplementedError):
        pj.solve(pj.POMCPSolver(estimate_value=pj.FOValue(np.zeros(4))), pj.LightDark1D())


@pytest.mark.parametrize("name", ["tiger", "baby", "rs44", "lightdark"])
def test_replay_tree_counts_bit_exact(oracle, gpu_lib, name):
    """north_star criterion 1: replayed uniform stream + injected rollout returns => identical counts."""
    pomdp, c = PROBS[name]
    gp, op = make_pair(oracle, pomdp, c, seed=0, mode="exact")
    rng = np.random.default_rng(123)
    Q = 800
    uniforms = rng.random(Q * (2 + 4 * 90))
    returns = rng.normal(-5.0, 10.0, Q)
    gp.set_belief_initial()
    op.set_belief_initial()
    a_g, N_g, V_g, uu_g, ru_g = gp.search_replay(Q, uniforms, returns)
    rc, a_o, N_o, V_o, uu_o, ru_o = op.search_replay(Q, uniforms, returns)
    assert rc == 0
    assert (uu_g, ru_g) == (uu_o, ru_o)
    assert np.array_equal(N_g, N_o)
    assert np.array_equal(_bits(V_g), _bits(V_o))
    assert a_g == a_o
    tg, to = gp.dump_tree(), op.dump_tree()
    assert np.array_equal(tg[0], to[0]) and np.array_equal(tg[1], to[1])
    gp.close()


@pytest.mark.parametrize("name", ["tiger", "baby", "rs44"])
def test_unweighted_update_episode_bit_exact(oracle, gpu_lib, name):
    """search -> re-root at (a,o) -> search ... : particles (in push order), counts and values agree."""
    pomdp, c = PROBS[name]
    gp, op = make_pair(oracle, pomdp, c, seed=21, mode="exact", tree_queries=1500)
    gp.set_belief_initial()
    op.set_belief_initial()
    for step in range(4):
        a_g, N_g, V_g = gp.search(1500)
        rc, a_o, N_o, V_o = op.search(1500)
        assert rc == 0 and a_g == a_o
        assert np.array_equal(N_g, N_o) and np.array_equal(_bits(V_g), _bits(V_o))
        # choose the most visited observation child of the chosen action (exists in both trees)
        best_o, best_n = None, 0
        for o in range(pomdp.n_observations):
            n, npart, _, _ = op.node_stats([(a_o, o)])
            if n > best_n:
                best_o, best_n = o, n
        assert best_o is not None
        ng, npg, Ng, Vg = gp.node_stats([(a_g, best_o)])
        no, npo, No, Vo = op.node_stats([(a_o, best_o)])
        assert (ng, npg) == (no, npo) and np.array_equal(Ng, No) and np.array_equal(_bits(Vg), _bits(Vo))
        gp.update_unweighted(a_g, best_o)
        assert op.update_unweighted(a_o, best_o) == 0
        pg, po = gp.belief_particles(), op.belief_particles()
        assert len(pg) == best_n == len(po)
        assert np.array_equal(pg, po)
    gp.close()


def test_unweighted_update_depletion(oracle, gpu_lib):
    """@test_throws ErrorException (test/Belief_and_Particle_Filter_Options.jl:35-36, test/runtests.jl:35-41)."""
    pomdp, c = PROBS["rs78"]
    gp, op = make_pair(oracle, pomdp, c, seed=2, mode="exact", tree_queries=5)
    gp.set_belief_initial()
    op.set_belief_initial()
    a, _, _ = gp.search(5)
    op.search(5)
    # an observation never generated under action `a` in 5 queries
    missing = None
    for o in range(3):
        if op.node_stats([(a, o)])[0] < 0:
            missing = o
            break
    assert missing is not None
    assert op.update_unweighted(a, missing) == abi.ERR_PARTICLE_DEPLETION
    with pytest.raises(pj.ParticleDepletion):
        gp.update_unweighted(a, missing)
    gp.close()


# ----------------------------------------------------------------------------- SIR update
@pytest.mark.parametrize("streaming", [False, True, "legacy"])
@pytest.mark.parametrize("name,n", [("baby", 10000), ("tiger", 4096), ("rs78", 50000), ("lightdark", 100000),
                                    ("lightdark", 1000000), ("rs78", 1 << 20)])
def test_sir_update_bit_exact(oracle, gpu_lib, name, n, streaming, monkeypatch):
    """fused kernel / tiled streaming kernels at the a-priori scale / max-scaled look-back kernels (legacy)"""
    if streaming:
        monkeypatch.setenv("POMCP_PF_STREAMING", "1")
    if streaming == "legacy":
        monkeypatch.setenv("POMCP_PF_LEGACY_STREAMING", "1")
    pomdp, c = PROBS[name]
    gp, op = make_pair(oracle, pomdp, c, seed=4, mode="exact", belief_mode=abi.BELIEF_EXTERNAL)
    rng = np.random.default_rng(9)
    states = random_states(pomdp, n, rng)
    gp.set_belief_particles(states)
    op.set_belief_particles(states)
    script = {"baby": [(0, 1), (1, 0), (0, 0)], "tiger": [(0, 1), (0, 1), (1, 0)],
              "rs78": [(5, 0), (0, 2), (7, 1)], "lightdark": [(2, 3.0), (2, 4.2), (0, 2.9)]}[name]
    for k, (a, o) in enumerate(script):
        n_out = n if k != 1 else n // 2 + 1
        gp.pf_update_sir(a, o, seed=1000 + k, n_out=n_out)
        assert op.pf_update_sir(a, o, seed=1000 + k, n_out=n_out, mode=0) == 0
        pg, po = gp.belief_particles(), op.belief_particles()
        assert len(pg) == n_out == len(po)
        assert pg.tobytes() == po.tobytes(), (name, k)
    gp.close()


@pytest.mark.parametrize("path", ["fused", "two_barriers", "streaming", "streaming_legacy"])
def test_sir_scale_rule_small_weights(oracle, gpu_lib, path, monkeypatch):
    """The fixed-point scale comes from the model's weight ceiling when that resolves the weights (one grid barrier
    on the device) and from the largest weight otherwise.  An observation far in the tail (o = 40: weights around
    1e-150) cannot be resolved at the a-priori scale: every path must fall back to the max-scaled integers and
    still equal the oracle bit for bit; a plausible observation right after uses the a-priori scale again."""
    if path.startswith("streaming"):  # tiled kernels at the a-priori scale (they hand over to the max-scaled ones)
        monkeypatch.setenv("POMCP_PF_STREAMING", "1")
    if path == "streaming_legacy":
        monkeypatch.setenv("POMCP_PF_LEGACY_STREAMING", "1")
    if path == "two_barriers":
        monkeypatch.setenv("POMCP_PF_TWO_BARRIERS", "1")
    ld = pj.LightDark1D()
    gp, op = make_pair(oracle, ld, 1.0, seed=4, mode="exact", belief_mode=abi.BELIEF_EXTERNAL)
    n = 30000
    rng = np.random.default_rng(21)
    parts = np.zeros(n, dtype=ld.state_dtype)
    parts["y"] = rng.normal(2.0, 3.0, n)
    gp.set_belief_particles(parts)
    op.set_belief_particles(parts)
    for k, (a, o) in enumerate([(2, 100.0), (2, 14.0), (2, 40.0), (0, 13.5)]):
        gp.pf_update_sir(a, o, seed=300 + k, n_out=n)
        assert op.pf_update_sir(a, o, seed=300 + k, n_out=n, mode=0) == 0
        assert gp.belief_particles().tobytes() == op.belief_particles().tobytes(), (path, k)
    gp.close()


@pytest.mark.parametrize("name,n,n_out", [("rs78", 2 * 2048 + 1, 5000), ("rs78", (1 << 21) + 17, (1 << 21) + 17),
                                          ("lightdark", 1_300_001, 1_000_000), ("baby", 1_250_000, 1_250_000),
                                          ("tiger", 6 * 2048 - 1, 2048),
                                          # more tiles than CTAs (8 per SM): a CTA's chunk holds several tiles, the last
                                          # chunks are short or empty, offsets are carried from tile to tile
                                          ("rs78", 1187 * 2048 + 5, 2_000_003), ("baby", 2_500_001, 2_500_001),
                                          ("rs78", 5000 * 2048 - 7, 5000 * 2048 - 7)])
def test_sir_tiled_path_edges_bit_exact(oracle, gpu_lib, name, n, n_out, monkeypatch):
    """The tiled kernels (weigh + tile / chunk sums, tile prefix where there is one, emit) at their edges: ragged last
    tiles, sets beyond one co-resident grid on the default path, chunks of several tiles per CTA, a block of terminal particles in FRONT (the first alive particle owns
    output 0 and sits in a later tile), every action class of the problem (RockSample: move / sample / sense, i.e. the
    weight-class table with and without a state change), n_out != n.  Bit-identical to the oracle's sequential walk."""
    if n < 1_200_000:
        monkeypatch.setenv("POMCP_PF_STREAMING", "1")  # small sets would take the fused kernel
    pomdp, c = PROBS[name]
    gp, op = make_pair(oracle, pomdp, c, seed=4, mode="exact", belief_mode=abi.BELIEF_EXTERNAL)
    rng = np.random.default_rng(31)
    states = random_states(pomdp, n, rng)
    lead = min(n // 3, 3000)
    if name == "rs78":
        states[:lead] |= np.uint32(abi.RS_TERMINAL)
    elif name == "lightdark":
        states["status"][:lead] = -1
    script = {"baby": [(0, 1), (1, 0)], "tiger": [(0, 1), (1, 0)], "rs78": [(5, 1), (1, 2), (4, 2), (9, 0)],
              "lightdark": [(2, 3.0), (0, 2.9)]}[name]
    for k, (a, o) in enumerate(script):
        gp.set_belief_particles(states)
        op.set_belief_particles(states)
        gp.pf_update_sir(a, o, seed=2000 + k, n_out=n_out)
        assert op.pf_update_sir(a, o, seed=2000 + k, n_out=n_out, mode=0) == 0
        pg, po = gp.belief_particles(), op.belief_particles()
        assert len(pg) == n_out == len(po)
        assert pg.tobytes() == po.tobytes(), (name, k, a, o)
    gp.close()


@pytest.mark.parametrize("seed", range(6))
def test_sir_tiled_path_random_cases_bit_exact(oracle, gpu_lib, seed):
    """Seeded random cases of the tiled path beyond one co-resident grid: set size (one to five tiles per CTA, ragged
    last tile), output size, problem, action, observation and the share of terminal particles are drawn; the
    resampled set is bit-identical to the oracle's sequential walk."""
    rng = np.random.default_rng(7100 + seed)
    name = ["rs78", "rs78", "baby", "tiger", "lightdark", "rs44"][seed]
    pomdp, c = PROBS[name]
    n = int(rng.integers(1_250_000, 11_000_000 if name != "lightdark" else 3_000_000))
    n_out = int(rng.integers(n // 3, 2 * n if name != "lightdark" else n))
    gp, op = make_pair(oracle, pomdp, c, seed=5, mode="exact", belief_mode=abi.BELIEF_EXTERNAL)
    states = random_states(pomdp, n, rng)
    for k in range(2):
        a = int(rng.integers(0, pomdp.n_actions))
        if name == "lightdark":
            o = float(rng.normal(3.0, 2.0))
        elif name in ("rs78", "rs44"):
            o = int(rng.integers(0, 2)) if a >= 5 else 2
        else:
            o = int(rng.integers(0, 2))
        gp.set_belief_particles(states)
        op.set_belief_particles(states)
        try:
            gp.pf_update_sir(a, o, seed=50 + k, n_out=n_out)
            rc_g = 0
        except pj.ParticleDepletion:
            rc_g = abi.ERR_PARTICLE_DEPLETION
        rc_o = op.pf_update_sir(a, o, seed=50 + k, n_out=n_out, mode=0)
        assert rc_g == rc_o, (name, n, n_out, a, o, rc_g, rc_o)
        if rc_o == 0:
            pg, po = gp.belief_particles(), op.belief_particles()
            assert len(pg) == n_out == len(po)
            assert pg.tobytes() == po.tobytes(), (name, n, n_out, a, o)
    gp.close()


def test_sir_baby_posterior_matches_golden_chain(oracle, gpu_lib):
    """D2: exact Baby belief chain 0 -> 0.0240964 -> 0.0298684 (notebooks/Display_Tree.ipynb:385):
    the 10^4-particle... here 2^20-particle SIR posterior mean must approach it within MC error."""
    pomdp = pj.BabyPOMDP(-5.0, -10.0)
    gp, _ = make_pair(oracle, pomdp, 10.0, seed=4, belief_mode=abi.BELIEF_EXTERNAL)
    n = 1 << 20
    gp.set_belief_particles(np.zeros(n, dtype=np.uint32))
    for k, want in enumerate([0.02409638554216867, 0.029868401596924436, 0.03128309542601844]):
        gp.pf_update_sir(0, 0, seed=50 + k, n_out=n)
        got = gp.belief_particles().mean()
        assert abs(got - want) < 5 * np.sqrt(want * (1 - want) / n) + 2e-4, (k, got, want)
    gp.close()


def test_sir_all_terminal(oracle, gpu_lib):
    pomdp = pj.LightDark1D()
    gp, op = make_pair(oracle, pomdp, 1.0, seed=4, belief_mode=abi.BELIEF_EXTERNAL)
    s = np.zeros(100, dtype=pomdp.state_dtype)
    s["status"] = -1
    gp.set_belief_particles(s)
    op.set_belief_particles(s)
    assert op.pf_update_sir(2, 3.0, 1, 100, 0) == abi.ERR_ALL_SAMPLES_TERMINAL
    with pytest.raises(pj.AllSamplesTerminal):
        gp.pf_update_sir(2, 3.0, 1, 100)
    with pytest.raises(pj.AllSamplesTerminal):
        gp.search(100)
    gp.close()


# ----------------------------------------------------------------------------- batched search
def _tree_invariants(node_N, act_N, child, root=0):
    """Q10: N(ha) = sum_o N(hao); N(h) = [expanded] + sum_a N(h,a) for visited h."""
    nn, nA = act_N.shape
    for n in range(nn):
        for a in range(nA):
            kids = child[n, a][child[n, a] >= 0]
            assert act_N[n, a] == node_N[kids].sum(), (n, a)


@pytest.mark.parametrize("name,queries", [("tiger", 20000), ("baby", 50000), ("rs78", 100000)])
def test_batched_search_invariants_and_values(oracle, gpu_lib, name, queries):
    """Batched mode (asynchronous workers, atomics): structural invariants hold exactly in every run;
    root statistics agree with the sequential oracle statistically.  Both solvers are randomised, so
    the comparison is between their distributions over 6 seeds.  Stated tolerance: the batched
    planner picks the oracle's modal root action in >= 5 of 6 runs, and the mean root value of that
    action differs by no more than 3 standard errors of the difference + 3% of the reward scale."""
    pomdp, c = PROBS[name]
    if name == "tiger":
        c = 1000.0  # random-rollout returns are O(-500): smaller c never revisits an action
    scale = {"tiger": 300.0, "baby": 15.0, "rs78": 10.0}[name]
    acts_g, acts_o, Vg, Vo = [], [], [], []
    for seed in range(6):
        gp, op = make_pair(oracle, pomdp, c, seed=31 + seed, mode="batched", tree_queries=queries)
        gp.set_belief_initial()
        op.set_belief_initial()
        a_g, N_g, V_g = gp.search(queries)
        rc, a_o, N_o, V_o = op.search(queries)
        assert rc == 0
        rootN, _, _ = gp.root_stats()
        assert rootN == queries
        assert N_g.sum() == queries - 1  # the root expansion consumes one query (src/solver.jl:104-106)
        if seed == 0:
            node_N, act_N, act_V, child = gp.dump_tree()
            _tree_invariants(node_N, act_N, child)
        cg = gp.counters()
        assert cg["simulations"] == queries and cg["rollouts"] <= 64 * (queries - 1)
        acts_g.append(a_g); acts_o.append(a_o); Vg.append(V_g); Vo.append(V_o)
        gp.close()
    modal = int(np.bincount(acts_o).argmax())
    assert sum(a == modal for a in acts_g) >= 5, (acts_g, acts_o)
    vg = np.array([v[modal] for v in Vg]); vo = np.array([v[modal] for v in Vo])
    se = np.sqrt(vg.var(ddof=1) / len(vg) + vo.var(ddof=1) / len(vo))
    assert abs(vg.mean() - vo.mean()) <= 3 * se + 0.03 * scale, (vg, vo)


@pytest.mark.parametrize("R", [1, 32, 64, 96, 128])
def test_batched_rollout_teams(oracle, gpu_lib, R):
    """rollouts_per_leaf = R in batched mode: a worker is a team of ceil(R/32) warps (4 above 64) that roll out
    each leaf together.  Every expanded non-root node gets exactly R rollouts, the counting invariants hold,
    and the root value of the chosen action agrees with the sequential oracle within 3 standard errors of
    the oracle's seed-to-seed spread + 3% of the reward scale (the leaf estimates only get less noisy with R)."""
    pomdp, c = PROBS["rs78"]
    Q = 60000
    vg, vo = [], []
    for seed in range(4):
        solver = pj.POMCPSolver(c=c, rng=71 + seed, tree_queries=Q, search_mode="batched", rollouts_per_leaf=R)
        gp = pj.solve(solver, pomdp)
        gp.set_belief_initial()
        a_g, N_g, V_g = gp.search(Q)
        cs = gp.counters()
        assert N_g.sum() == Q - 1 and cs["simulations"] == Q
        assert cs["rollouts"] % R == 0 and cs["rollouts"] // R <= cs["nodes"] - 1
        assert cs["rollout_steps"] <= cs["rollouts"] * 89
        if seed == 0:
            node_N, act_N, act_V, child = gp.dump_tree()
            kid_N = np.where(child >= 0, node_N[np.maximum(child, 0)], 0).sum(axis=2)
            assert np.array_equal(kid_N, act_N)
        gp.close()
        sp = solver.c_params(abi.BELIEF_UNWEIGHTED)
        sp.search_mode, sp.rollouts_per_leaf = abi.MODE_EXACT, 1
        op2 = oracle.OraclePlanner(pomdp, sp)
        op2.set_belief_initial()
        rc, a_o, N_o, V_o = op2.search(Q)
        assert rc == 0 and a_g == a_o
        vg.append(V_g[a_o]); vo.append(V_o[a_o])
    se = np.std(vo, ddof=1) / np.sqrt(len(vo)) + np.std(vg, ddof=1) / np.sqrt(len(vg))
    assert abs(np.mean(vg) - np.mean(vo)) <= 3 * se + 0.3, (vg, vo)


def test_batched_lightdark_continuous_obs(oracle, gpu_lib):
    pomdp, c = PROBS["lightdark"]
    gp, op = make_pair(oracle, pomdp, c, seed=8, mode="batched", tree_queries=30000)
    gp.set_belief_initial()
    op.set_belief_initial()
    a_g, N_g, V_g = gp.search(30000)
    rc, a_o, N_o, V_o = op.search(30000)
    assert N_g.sum() == 29999
    cg, co = gp.counters(), op.counters()
    assert cg["max_depth_seen"] == co["max_depth_seen"] == 1  # Float64 observations never repeat
    assert cg["nodes"] == 30000
    assert abs(V_g[N_g.argmax()] - V_o[N_o.argmax()]) < 1.0
    gp.close()


def test_batched_particle_belief_and_update(oracle, gpu_lib):
    """Batched search from a particle belief, then the unweighted re-root: the new belief holds
    exactly N(child) states, all consistent with the model."""
    pomdp, c = PROBS["baby"]
    gp, _ = make_pair(oracle, pomdp, c, seed=5, mode="batched", tree_queries=20000)
    rng = np.random.default_rng(3)
    gp.set_belief_particles((rng.random(10000) < 0.1).astype(np.uint32))
    a, N, V = gp.search(20000)
    n_child, npart, _, _ = gp.node_stats([(a, 0)])
    assert n_child > 0
    gp.update_unweighted(a, 0)
    p = gp.belief_particles()
    assert len(p) == n_child
    assert set(np.unique(p)) <= {0, 1}
    rootN, _, _ = gp.root_stats()
    assert rootN == n_child
    a2, N2, V2 = gp.search(20000)
    rootN2, _, _ = gp.root_stats()
    assert rootN2 == n_child + 20000  # tree reuse: counts continue (src/solver.jl:276)
    gp.close()


# ----------------------------------------------------------------------------- API surface
def test_api_episode_baby(gpu_lib):
    """Basic_Usage notebook flow through the mirrored API: solve / action / updater / update."""
    pomdp = pj.BabyPOMDP(-5.0, -10.0)
    solver = pj.POMCPSolver(tree_queries=3000, c=10.0, rng=7)
    policy = pj.solve(solver, pomdp)
    up = pj.updater(policy)
    b = pj.initialize_belief(up, pj.initial_state_distribution(pomdp))
    a = pj.action(policy, b)
    assert a in (0, 1)
    assert b.N == 3000
    kids = b.children
    assert sum(n for n, _ in kids.values()) == 2999
    b2 = pj.update(up, b, a, 0)
    assert len(b2.particles()) == b2.N > 0
    with pytest.raises(ValueError):
        _ = b.N  # stale node
    a2 = pj.action(policy, b2)
    assert a2 in (0, 1)
    policy.close()


def test_api_default_action_and_errors(gpu_lib):
    pomdp = pj.LightDark1D()
    s = np.zeros(10, dtype=pomdp.state_dtype)
    s["status"] = -1
    policy = pj.solve(pj.POMCPSolver(tree_queries=50, rng=1), pomdp)
    with pytest.raises(pj.AllSamplesTerminal):
        pj.action(policy, pj.ParticleCollection(s))
    policy.close()
    policy = pj.solve(pj.POMCPSolver(tree_queries=50, rng=1, default_action=2), pomdp)
    assert pj.action(policy, pj.ParticleCollection(s)) == 2
    policy.close()
    with pytest.raises(ValueError):
        pj.solve(pj.POMCPSolver(), pj.BabyPOMDP(discount_factor=1.0))
    with pytest.raises(NotImplementedError):
        pj.solve(pj.POMCPSolver(num_sparse_actions=3), pj.BabyPOMDP())


# ----------------------------------------------------------------------------- episode reward (criterion 3)
def _run_episodes(oracle, pomdp, make_planner, n_episodes, max_steps, seed0, use_gpu, n_part=0, first_ep=0):
    """RolloutSimulator outer loop: action -> environment step -> belief update.
    The environment is stepped by the oracle's generative model with a host RNG shared by both arms."""
    rewards = []
    for ep in range(first_ep, first_ep + n_episodes):
        env = np.random.default_rng(seed0 + ep)
        pl = make_planner(ep)
        if n_part:
            s0 = pl_init_particles(pomdp, n_part, env)
            pl.set_belief_particles(s0)
            s = s0[0].copy()
        else:
            pl.set_belief_initial()
            s = np.zeros(1, dtype=pomdp.state_dtype)[0]
        disc, total = 1.0, 0.0
        for t in range(max_steps):
            if use_gpu:
                try:
                    a = pl.search(-1)[0]
                except pj.AllSamplesTerminal:
                    break
            else:
                rc, a, _, _ = pl.search(-1)
                if rc == abi.ERR_ALL_SAMPLES_TERMINAL:
                    break
                assert rc == 0
            if oracle.lib().orc_is_terminal(pomdp.c_struct(), np.array([s]).ctypes.data):
                break
            s, o, r = oracle.model_step(pomdp, s, a, ut=(env.random(), 0), uo=(env.random(), env.random()))
            total += disc * r
            disc *= pomdp.discount
            o = pj.obs_from_bits(pomdp, o)
            try:
                if n_part:
                    if use_gpu:
                        pl.pf_update_sir(a, o, seed=ep * 1000 + t, n_out=n_part)
                    else:
                        rc = pl.pf_update_sir(a, o, seed=ep * 1000 + t, n_out=n_part, mode=0)
                        if rc != 0:
                            break
                else:
                    if use_gpu:
                        pl.update_unweighted(a, o)
                    elif pl.update_unweighted(a, o) != 0:
                        total = None
                        break
            except (pj.ParticleDepletion, pj.AllSamplesTerminal):
                if not n_part:
                    total = None
                break
        if total is not None:
            rewards.append(total)
        pl.close()
    return np.array(rewards)


def pl_init_particles(pomdp, n, rng):
    rocks = rng.integers(0, 1 << pomdp.k, n).astype(np.uint32)
    return (np.uint32(pomdp.start[0]) | np.uint32(pomdp.start[1] << 4) | (rocks << 8)).astype(np.uint32)


def test_episode_reward_baby_not_worse(oracle, gpu_lib):
    """north_star criterion 3 on the reference's own sanity setting (Baby, tree_queries=300, c=10,
    tree reuse through the RootUpdater): the batched GPU planner's mean discounted reward must not be
    worse than the sequential oracle's (one-sided, 3 standard errors of the difference)."""
    pomdp = pj.BabyPOMDP(-5.0, -10.0)

    def gpu(ep):
        return pj.solve(pj.POMCPSolver(c=10.0, rng=500 + ep, tree_queries=300, search_mode="batched"), pomdp)

    def cpu(ep):
        return oracle.OraclePlanner(pomdp, oracle.default_params(c=10.0, seed=500 + ep, tree_queries=300))

    rg = _run_episodes(oracle, pomdp, gpu, 80, 44, 9000, True)
    ro = _run_episodes(oracle, pomdp, cpu, 80, 44, 9000, False)
    assert len(rg) >= 70 and len(ro) >= 70
    se = np.sqrt(rg.var() / len(rg) + ro.var() / len(ro))
    assert rg.mean() >= ro.mean() - 3 * se, (rg.mean(), ro.mean(), se)
    assert -22.0 < rg.mean() < -13.0  # reference band (notebooks/Sanity_Checks.ipynb:197,239,134)


def test_episode_reward_rocksample_not_worse(oracle, gpu_lib):
    """Same criterion on RockSample(4,4) with the SIR belief (fresh root every step)."""
    pomdp = pj.RockSamplePOMDP(4, 4, seed=44)
    Q, n_part = 3000, 2000

    def gpu(ep):
        return pj.solve(pj.POMCPSolver(c=20.0, rng=700 + ep, tree_queries=Q, search_mode="batched"), pomdp,
                        belief_mode=abi.BELIEF_EXTERNAL)

    def cpu(ep):
        return oracle.OraclePlanner(pomdp, oracle.default_params(c=20.0, seed=700 + ep, tree_queries=Q,
                                                                 belief_mode=abi.BELIEF_EXTERNAL))

    rg = _run_episodes(oracle, pomdp, gpu, 40, 30, 9500, True, n_part)
    ro = _run_episodes(oracle, pomdp, cpu, 40, 30, 9500, False, n_part)
    se = np.sqrt(rg.var() / len(rg) + ro.var() / len(ro))
    assert rg.mean() >= ro.mean() - 3 * se, (rg.mean(), ro.mean(), se)
    assert rg.mean() > 0.0  # a useful policy: exits and samples good rocks


def test_rocksample_11_11_exact_and_batched(oracle, gpu_lib):
    """BASELINE config 5 problem (16 actions): exact parity and batched agreement on one GPU."""
    pomdp = pj.RockSamplePOMDP(11, 11, seed=11011)
    gp, op = make_pair(oracle, pomdp, 20.0, seed=3, mode="exact", tree_queries=3000)
    gp.set_belief_initial()
    op.set_belief_initial()
    a_g, N_g, V_g = gp.search(3000)
    rc, a_o, N_o, V_o = op.search(3000)
    assert rc == 0 and a_g == a_o and np.array_equal(N_g, N_o) and np.array_equal(_bits(V_g), _bits(V_o))
    gp.close()
    gp, op = make_pair(oracle, pomdp, 20.0, seed=3, mode="batched", tree_queries=60000)
    gp.set_belief_initial()
    op.set_belief_initial()
    a_g, N_g, V_g = gp.search(60000)
    rc, a_o, N_o, V_o = op.search(60000)
    assert N_g.sum() == 59999
    assert V_g[a_o] >= V_g.max() - 0.5 and abs(V_g.max() - V_o.max()) < 0.15 * abs(V_o.max()) + 0.3, (V_g, V_o)
    gp.close()

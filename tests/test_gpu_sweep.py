"""Randomised parameter sweep (exact mode, bit parity with the oracle) and the edge cases of the reference's
solver fields (src/POMCP.jl:75-136, src/constructor.jl:8-18): eps, max_depth, c (including 0: the NaN
criterion of src/solver.jl:117-119), init_V (including -Inf: src/solver.jl:114-115,152), init_N,
rollouts_per_leaf, estimator, discount, tree reuse across several decisions."""
import numpy as np
import pytest

import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

from helpers import make_pair

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


def _same(gp, op, Q):
    a_g, N_g, V_g = gp.search(Q)
    rc, a_o, N_o, V_o = op.search(Q)
    assert rc == 0
    assert a_g == a_o and np.array_equal(N_g, N_o), (N_g, N_o)
    assert np.array_equal(_bits(V_g), _bits(V_o)), (V_g, V_o)
    cg, co = gp.counters(), op.counters()
    for k in ("simulations", "terminal_skips", "rollouts", "rollout_steps", "tree_steps", "nodes", "log_entries",
              "max_depth_seen"):
        assert cg[k] == co[k], (k, cg[k], co[k])
    return a_g


def test_randomised_solver_parameters_bit_exact(oracle, gpu_lib):
    rng = np.random.default_rng(2026)
    makers = [lambda g: pj.TigerPOMDP(discount_factor=g), lambda g: pj.BabyPOMDP(-5.0, -10.0, discount_factor=g),
              lambda g: pj.RockSamplePOMDP(4, 4, seed=44, discount_factor=g), lambda g: pj.LightDark1D(discount_factor=g),
              lambda g: pj.RockSamplePOMDP(7, 8, seed=7008, discount_factor=g)]
    for trial in range(40):
        gamma = float(rng.choice([0.5, 0.8, 0.9, 0.95, 0.99]))
        pomdp = makers[trial % len(makers)](gamma)
        kw = dict(
            eps=float(rng.choice([0.3, 0.05, 0.01, 0.001])),
            max_depth=int(rng.choice([1, 2, 5, 20, 2 ** 62])),
            c=float(rng.choice([0.0, 0.5, 1.0, 10.0, 100.0])),
            init_V=float(rng.choice([0.0, -3.0, 5.0, -np.inf])),
            init_N=int(rng.choice([0, 0, 1, 3])),
            rollouts_per_leaf=int(rng.choice([1, 1, 2, 7, 32, 33, 64])),
        )
        est = rng.integers(0, 3)
        if est == 1:
            kw["estimate_value"] = float(rng.normal(0, 5))
        elif est == 2 and isinstance(pomdp, pj.BabyPOMDP):
            kw["estimate_value"] = pj.RolloutEstimator(pj.FeedWhenCrying())
        Q = int(rng.integers(1, 700))
        gp, op = make_pair(oracle, pomdp, kw.pop("c"), seed=int(rng.integers(0, 2 ** 40)), mode="exact", tree_queries=Q, **kw)
        gp.set_belief_initial()
        op.set_belief_initial()
        try:
            _same(gp, op, Q)
            _same(gp, op, max(1, Q // 3))  # tree reuse, next epoch
        except AssertionError:
            print("trial", trial, type(pomdp).__name__, gamma, kw, Q)
            raise
        gp.close()


def test_zero_queries_and_all_terminal(oracle, gpu_lib):
    """tree_queries = 0 throws AllSamplesTerminal like an all-terminal belief (src/solver.jl:56-58)."""
    pomdp = pj.BabyPOMDP(-5.0, -10.0)
    gp, op = make_pair(oracle, pomdp, 10.0, seed=1, mode="exact", tree_queries=10)
    gp.set_belief_initial()
    op.set_belief_initial()
    with pytest.raises(pj.AllSamplesTerminal):
        gp.search(0)
    assert op.search(0)[0] == abi.ERR_ALL_SAMPLES_TERMINAL
    gp.close()
    rs = pj.RockSamplePOMDP(4, 4, seed=44)
    for mode in ("exact", "batched"):
        gp, _ = make_pair(oracle, rs, 20.0, seed=1, mode=mode, tree_queries=100)
        gp.set_belief_particles(np.full(50, abi.RS_TERMINAL, dtype=np.uint32))
        with pytest.raises(pj.AllSamplesTerminal):
            gp.search(100)
        assert gp.counters()["terminal_skips"] == 100 and gp.counters()["simulations"] == 0
        gp.close()


def test_mixed_terminal_belief_skips_are_counted(oracle, gpu_lib):
    """Terminal root samples consume their query without a simulation or an N increment (src/solver.jl:49-53):
    Root.N = number of non-terminal draws, in exact (bit parity) and batched mode."""
    rs = pj.RockSamplePOMDP(4, 4, seed=44)
    rng = np.random.default_rng(3)
    parts = (rng.integers(0, 16, 400).astype(np.uint32) << 8) | np.uint32(1 | (2 << 4))
    parts[::3] |= np.uint32(abi.RS_TERMINAL)
    gp, op = make_pair(oracle, rs, 20.0, seed=5, mode="exact", tree_queries=3000)
    gp.set_belief_particles(parts)
    op.set_belief_particles(parts)
    _same(gp, op, 3000)
    cs = gp.counters()
    assert cs["terminal_skips"] > 800 and gp.root_stats()[0] == 3000 - cs["terminal_skips"]
    gp.close()
    gp, _ = make_pair(oracle, rs, 20.0, seed=5, mode="batched", tree_queries=30000)
    gp.set_belief_particles(parts)
    a, N, V = gp.search(30000)
    cs = gp.counters()
    assert cs["simulations"] + cs["terminal_skips"] == 30000
    assert gp.root_stats()[0] == cs["simulations"] and N.sum() == cs["simulations"] - 1
    gp.close()


@pytest.mark.gpu
def test_inflight_cap_follows_the_budget(gpu_lib):
    """Simulations in flight (pomcp_counters::waves reports the admission cap of the last search, per tree): left to the
    library it is, on RockSample, min(16 one-warp workers per SM, max(64, tree_queries / 256)) - the budget-relative cap
    was measured to matter there - and 8 two-warp teams per SM elsewhere; an explicit inflight_max is taken as given (clamped to what is co-resident);
    several trees share the workers."""
    import torch
    sms = torch.cuda.get_device_properties(0).multi_processor_count
    rs = pj.RockSamplePOMDP(7, 8, seed=7008)

    def cap(pomdp, q, **kw):
        pl = pj.solve(pj.POMCPSolver(c=20.0, rng=1, tree_queries=q, search_mode="batched", **kw), pomdp)
        pl.set_belief_initial()
        pl.search(q)
        w = pl.counters()["waves"]
        pl.close()
        return w

    assert cap(rs, 100_000) == 390
    assert cap(rs, 1_000) == 64
    assert cap(rs, 1_000_000) == 16 * sms
    assert cap(rs, 100_000, inflight_max=8 * sms) == 8 * sms
    assert cap(rs, 100_000, inflight_max=100_000) == 16 * sms
    assert cap(pj.BabyPOMDP(-5.0, -10.0), 1_000_000) == 8 * sms
    assert cap(pj.BabyPOMDP(-5.0, -10.0), 40_000) == 8 * sms
    assert cap(rs, 1_000_000, n_trees=8) == 2 * sms

"""GPU tests of the outer simulation loop: RolloutSimulator + simulate +
test_solver on the device model (pomcp_generate_sor), the reference's own smoke tests
(test/runtests.jl:13-31) re-expressed, and the create_json tree export (src/visualization.jl:12-72)."""
import json

import numpy as np
import pytest

import pomcp_jl_b200 as pj

pytestmark = pytest.mark.gpu


def test_generate_sor_matches_oracle_model(oracle, gpu_lib):
    """The environment step uses the same model code as the planner: with the same uniforms it equals the
    oracle's model_step for every problem."""
    for pomdp in (pj.TigerPOMDP(), pj.BabyPOMDP(-5.0, -10.0), pj.RockSamplePOMDP(4, 4, seed=44), pj.LightDark1D()):
        pl = pj.solve(pj.POMCPSolver(tree_queries=10), pomdp)
        s = pl.initial_state(seed=3)
        for step in range(1, 30):
            a = step % pomdp.n_actions
            w = pj.philox4x32_10((0, 6, step, 0), (77, 0))
            sp, o, r, term = pl.generate_sor(s, a, 77, step)
            sp_o, o_bits, r_o = oracle.model_step(pomdp, s, a, ut=(w[0] * 2.0 ** -32, 0.0),
                                                  uo=(w[2] * 2.0 ** -32, w[3] * 2.0 ** -32))
            assert np.array([sp]).tobytes() == np.array([sp_o]).tobytes() and r == r_o
            assert pj.obs_bits(o) == o_bits
            if term:
                break
            s = sp
        pl.close()


def test_reference_smoke_tests(gpu_lib):
    """test/runtests.jl:13-31: the three POMCPSolver smoke tests run an episode without error."""
    for est in (pj.RolloutEstimator(pj.FeedWhenCrying()), None, pj.FORollout(pj.FeedWhenCrying())):
        kw = {} if est is None else {"estimate_value": est, "eps": 0.01, "c": 10.0, "tree_queries": 50, "rng": 2}
        solver = pj.POMCPSolver(**kw)
        r = pj.test_solver(solver, pj.BabyPOMDP(-5.0, -10.0))
        assert np.isfinite(r) and -150.0 <= r <= 0.0
    # particle depletion contract (test/runtests.jl:35-41)
    solver = pj.POMCPSolver(estimate_value=pj.RolloutEstimator(pj.FeedWhenCrying()), eps=0.01, c=10.0, tree_queries=5, rng=2)
    with pytest.raises(pj.ParticleDepletion):
        pj.test_solver(solver, pj.BabyPOMDP(-5.0, -10.0), max_steps=100)


def test_simulate_with_sir_updater_and_rocksample(gpu_lib):
    """simulate() with an external SIR updater (test/runtests.jl:57 pattern) on LightDark1D, and a full
    RockSample(7,8) episode with the unweighted updater: rewards are finite, histories consistent."""
    ld = pj.LightDark1D()
    policy = pj.solve(pj.POMCPSolver(tree_queries=2000, c=10.0, rng=4), ld, belief_mode=pj._abi.BELIEF_EXTERNAL)
    rng = np.random.default_rng(0)
    parts = np.zeros(500, dtype=ld.state_dtype)
    parts["y"] = rng.normal(2.0, 3.0, 500)
    sim = pj.RolloutSimulator(rng=5, max_steps=15)
    r = pj.simulate(sim, ld, policy, pj.SIRParticleFilter(policy, 500, rng=1), pj.ParticleCollection(parts))
    assert np.isfinite(r) and len(sim.history) >= 1
    policy.close()
    rs = pj.RockSamplePOMDP(7, 8, seed=7008)
    policy = pj.solve(pj.POMCPSolver(tree_queries=20000, c=20.0, rng=6,
                                     node_belief_updater=pj.SIRReinvigorator(2000, rng=3)), rs)
    sim = pj.RolloutSimulator(rng=11, max_steps=40)
    r = pj.simulate(sim, rs, policy)
    assert np.isfinite(r) and r >= -10.0
    assert abs(sum(0.95 ** t * h[3] for t, h in enumerate(sim.history)) - r) < 1e-9
    policy.close()


def test_create_json_tree_export(gpu_lib):
    """create_json: ids are 1..n in depth-first order, children_ids are consistent, N(h) = [expanded] + sum N(ha)."""
    baby = pj.BabyPOMDP(-5.0, -10.0)
    policy = pj.solve(pj.POMCPSolver(tree_queries=200, c=10.0, rng=1, search_mode="exact"), baby)
    b = pj.initialize_belief(pj.updater(policy), pj.initial_state_distribution(baby))
    pj.action(policy, b)
    js, root_id = pj.create_json(policy)
    nd = {int(k): v for k, v in json.loads(js).items()}
    assert root_id == 1 and nd[1]["type"] == "belief" and nd[1]["N"] == 200
    assert sorted(nd) == list(range(1, len(nd) + 1))
    for i, n in nd.items():
        assert all(c in nd and c > i for c in n["children_ids"])
        if n["type"] == "belief" and n["children_ids"]:
            assert n["N"] == 1 + sum(nd[c]["N"] for c in n["children_ids"])
        if n["type"] == "action":
            assert n["N"] == sum(nd[c]["N"] for c in n["children_ids"])
    policy.close()


def test_tree_navigation_like_the_reference(oracle, gpu_lib):
    """b.children[a].children[o] walks the device tree like the reference's node objects (src/tree.jl:7-26; the
    notebooks' `delete!(b.children[a].children, o)` idiom reads the same fields): N(h) = 1 + sum_a N(ha),
    N(ha) = sum_o N(hao), and the values equal the oracle's for the same history (exact mode)."""
    baby = pj.BabyPOMDP(-5.0, -10.0)
    solver = pj.POMCPSolver(tree_queries=500, c=10.0, rng=3, search_mode="exact")
    policy = pj.solve(solver, baby)
    b = pj.initialize_belief(pj.updater(policy), pj.initial_state_distribution(baby))
    a = pj.action(policy, b)
    op = oracle.OraclePlanner(baby, solver.c_params(pj._abi.BELIEF_UNWEIGHTED))
    op.set_belief_initial()
    assert op.search(500)[1] == a
    assert isinstance(b, pj.RootNode) and b.N == 500 and set(b.children) == {0, 1}
    assert b.N == 1 + sum(n.N for n in b.children.values())
    for act, an in b.children.items():
        assert isinstance(an, pj.ActNode) and an.label == act
        assert an.N == sum(c.N for c in an.children.values())
        for o, hao in an.children.items():
            assert isinstance(hao, pj.ObsNode) and hao.label == o
            nN, npart, N_o, V_o = op.node_stats([(act, o)])
            assert hao.N == nN
            for a2, an2 in hao.children.items():
                assert an2.N == N_o[a2] and an2.V == V_o[a2]
    assert 5 not in b.children[0].children and b.children[0].children.get(5) is None
    pj.update(pj.updater(policy), b, a, 0)
    with pytest.raises(ValueError):
        b.N  # stale after the re-root, like a dropped Julia reference
    assert pj.convert_estimator(pj.PORollout(pj.RandomSolver(), None), solver, baby) == pj._abi.EST_RANDOM_ROLLOUT
    assert pj.sparse_actions(policy, baby) == [0, 1] and pj.init_V(3, baby) == 3.0 and pj.init_N(2.0, baby) == 2
    # widening is active on a discrete space larger than k_observation (tests/test_gpu_dpw.py has the parity tests)
    widened = pj.solve(pj.POMCPDPWSolver(tree_queries=100, enable_action_pw=False, k_observation=1.0), baby)
    assert widened._sp.dpw_observation == 1
    widened.close()
    policy.close()
    # test/runtests.jl:49-54: DPW without action widening on Baby never widens (2 observations <= k_observation), but it
    # is still POMCPDPWSolver's simulate(): its leaf rollouts get `depth` as their step budget (src/solver.jl:199)
    # where POMCPSolver's run to the horizon (src/solver.jl:97-102).  Bit-identical to the oracle with dpw_solver = 1,
    # and not the tree POMCPSolver builds with the same parameters.
    dpw = pj.POMCPDPWSolver(tree_queries=100, eps=0.01, c=10.0, enable_action_pw=False, rng=2, search_mode="exact")
    r = pj.test_solver(dpw, baby)
    assert np.isfinite(r)
    p1, p2 = pj.solve(dpw, baby), pj.solve(pj.POMCPSolver(tree_queries=100, eps=0.01, c=10.0, rng=2, search_mode="exact"), baby)
    op = oracle.OraclePlanner(baby, oracle.default_params(tree_queries=100, eps=0.01, c=10.0, seed=2, dpw_solver=1))
    for p_ in (p1, p2, op):
        p_.set_belief_initial()
    r1, r2 = p1.search(100), p2.search(100)
    rc, a_o, N_o, V_o = op.search(100)
    assert rc == 0 and r1[0] == a_o and np.array_equal(r1[1], N_o) and np.array_equal(r1[2].view(np.uint64), V_o.view(np.uint64))
    assert p1.counters()["rollout_steps"] < p2.counters()["rollout_steps"] / 3
    p1.close(); p2.close()


def test_episode_reward_rocksample_7_8_not_worse_than_oracle(oracle, gpu_lib):
    """north_star criterion 3 at the headline problem: RockSample(7,8), SIR belief of 2^16 particles, 292 seeded
    40-step episodes, 10^5 simulations per decision, the batched planner (library defaults) against the sequential
    oracle on the same planner seeds and the same environment draws.  The oracle's rewards are a committed fixture
    (tests/golden/episodes_rs78_oracle_1e5_292.json, written by scripts/episode_table.py --arm oracle: the oracle is
    deterministic; its first episode is re-run here to make sure the fixture still belongs to this oracle).
    One-sided paired test: mean(batched - oracle) >= -3 standard errors of the paired difference (0.16 of 15.7; round 2
    measured +0.09 +- 0.16 with the budget-relative in-flight cap and -0.81 +- 0.18 with 2368 in flight, which this
    test rejects)."""
    import json
    import os
    import sys
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "scripts"))
    import episode_table as et
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    small = json.load(open(os.path.join(gold, "episodes_rs78_oracle_1e4.json")))
    assert et.one_episode(("oracle", 10000, 0, {}))[1] == small["rewards"][0]  # the fixture is this oracle's
    ref = json.load(open(os.path.join(gold, "episodes_rs78_oracle_1e5_292.json")))
    ro = np.array(ref["rewards"])
    rg = np.array([et.one_episode(("gpu", 100000, ep, {}))[1] for ep in range(len(ro))])
    d = rg - ro
    se = d.std(ddof=1) / np.sqrt(len(d))
    print(f"episodes at 1e5: batched {rg.mean():.2f}, oracle {ro.mean():.2f}, paired difference {d.mean():.2f} +- {se:.2f}")
    assert d.mean() >= -3 * se - 1e-9, (rg.mean(), ro.mean(), se)
    # the 24-episode fixture of round 2's first sessions is the head of the 292-episode one
    assert json.load(open(os.path.join(gold, "episodes_rs78_oracle_1e5.json")))["rewards"] == ref["rewards"][:24]
    assert rg.mean() > 10.0  # a useful policy: samples good rocks and leaves

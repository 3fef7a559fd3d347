"""GPU tests of the generic node-belief-updater branch (src/solver.jl:138, src/updater.jl:30-39) with the problem's
exact two-state updater - the configuration of notebooks/Sanity_Checks.ipynb:47,70 (`node_belief_updater =
updater(problem)`) and test/Belief_and_Particle_Filter_Options.jl:63-77: every node carries P(s = 1), the root
state is drawn from it, a new node's belief is update(updater, h.B, a, o), and update() never depletes."""
import numpy as np
import pytest

import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

from helpers import make_pair

pytestmark = pytest.mark.gpu


def _bits(a):
    return np.asarray(a, dtype=np.float64).view(np.uint64)


@pytest.mark.parametrize("name,c,p0,R", [("baby", 10.0, 0.0, 1), ("baby", 10.0, 0.4, 4), ("tiger", 100.0, 0.5, 1)])
def test_exact2_episode_bit_exact(oracle, gpu_lib, name, c, p0, R):
    """Exact mode: decisions, root statistics and the root belief equal the oracle's over a 25-step episode, bit for
    bit, through re-roots onto existing children and onto children the tree does not hold."""
    pomdp = pj.BabyPOMDP(-5.0, -10.0) if name == "baby" else pj.TigerPOMDP()
    fresh = 0
    for seed in (3, 4):
        gp, op = make_pair(oracle, pomdp, c, seed=seed, mode="exact", belief_mode=abi.BELIEF_EXACT2, tree_queries=400,
                           rollouts_per_leaf=R)
        gp.set_belief_exact2(p0)
        assert op.set_belief_exact2(p0) == 0
        rng = np.random.default_rng(seed)
        s, b = np.uint32(rng.random() < p0), p0
        for t in range(25):
            Q = 400 if t >= 6 else 3  # tiny searches first: the observed child is then often missing from the tree
            a_g, N_g, V_g = gp.search(Q)
            rc, a_o, N_o, V_o = op.search(Q)
            assert rc == 0 and a_g == a_o and np.array_equal(N_g, N_o) and np.array_equal(_bits(V_g), _bits(V_o)), (t, N_g, N_o)
            s, o, r = oracle.model_step(pomdp, s, a_g, ut=(rng.random(), 0), uo=(rng.random(), 0))
            gp.update_exact2(a_g, int(o))
            assert op.update_exact2(a_g, int(o)) == 0
            b = oracle.bayes2(pomdp, b, a_g, int(o))
            assert gp.belief_exact2() == op.belief_exact2() == b
            fresh += gp.root_stats()[0] == 0
            assert gp.root_stats()[0] == op.root_stats()[0]
        cg, co = gp.counters(), op.counters()
        for key in ("simulations", "rollouts", "rollout_steps", "tree_steps"):
            assert cg[key] == co[key], (key, cg[key], co[key])
        gp.close()
    assert fresh >= 1  # both branches of src/updater.jl:30-39 were taken


def test_exact2_batched_and_api(oracle, gpu_lib):
    """Batched mode keeps the counting invariants and the exact root belief; through the API the planner is built
    the way the notebook builds it, and an episode of 44 steps never depletes (test/runtests.jl:35-41 does, with
    the particle updater)."""
    baby = pj.BabyPOMDP(-5.0, -10.0)
    solver = pj.POMCPSolver(eps=0.01, c=10.0, tree_queries=300, rng=2, node_belief_updater=pj.updater(baby))
    policy = pj.solve(solver, baby)
    assert policy.belief_mode == abi.BELIEF_EXACT2
    up = pj.updater(policy)
    b = pj.initialize_belief(up, pj.BoolDistribution(0.0))
    a = pj.action(policy, b)
    rn, N, V = policy.root_stats()
    assert rn == 300 and N.sum() == 299 and a in (0, 1)
    bel, s = 0.0, np.uint32(0)
    rng = np.random.default_rng(0)
    for t in range(44):
        a = pj.action(policy, b)
        s, o, r = oracle.model_step(baby, s, a, ut=(rng.random(), 0), uo=(rng.random(), 0))
        b = pj.update(up, b, a, int(o))
        bel = oracle.bayes2(baby, bel, a, int(o))
        assert b.B.p == bel == policy.belief_exact2()
    policy.close()
    # the whole loop through simulate(), initial belief = the distribution itself
    policy = pj.solve(pj.POMCPSolver(eps=0.01, c=10.0, tree_queries=300, rng=3, node_belief_updater=pj.updater(baby)), baby)
    sim = pj.RolloutSimulator(rng=7, max_steps=44)
    r = pj.simulate(sim, baby, policy, pj.updater(policy), pj.initial_state_distribution(baby))
    assert len(sim.history) == 44 and -80.0 < r < 0.0
    policy.close()
    # RockSample has no exact device updater
    with pytest.raises(NotImplementedError):
        pj.solve(pj.POMCPSolver(node_belief_updater=pj.updater(pj.RockSamplePOMDP(4, 4, seed=1))), pj.RockSamplePOMDP(4, 4, seed=1))

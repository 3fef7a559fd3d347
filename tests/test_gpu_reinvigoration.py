"""GPU tests of the device-side reinvigorator: update(::RootUpdater{R}, b, a, o)
with R = SIRReinvigorator(n_min), i.e. the reinvigorate! / handle_unseen_observation hooks of
src/particle_filter.jl:37-73 evaluated on the device from the SIR posterior of the old root belief.
Exact mode is compared with the oracle bit for bit (decisions, outcomes, particle sets)."""
import numpy as np
import pytest

import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi

from helpers import make_pair, problems

pytestmark = pytest.mark.gpu
PROBS = problems()


@pytest.mark.parametrize("name,queries,n_min,steps", [("baby", 5, 100, 60), ("tiger", 30, 64, 30), ("rs44", 200, 256, 12),
                                                      ("lightdark", 50, 128, 8)])
def test_reinvigorated_episode_bit_exact(oracle, gpu_lib, name, queries, n_min, steps, monkeypatch):
    """Low query counts deplete the default updater within a few steps (the reference's own test expects
    the exception, test/runtests.jl:35-41); with the reinvigorator the episode runs on, and every
    decision, outcome code and belief particle set equals the oracle's."""
    monkeypatch.setenv("POMCP_COMPACT_ALWAYS", "1")  # also exercises compaction after a top-up
    pomdp, c = PROBS[name]
    gp, op = make_pair(oracle, pomdp, c, seed=13, mode="exact", tree_queries=queries)
    gp.set_belief_initial()
    op.set_belief_initial()
    rng = np.random.default_rng(5)
    s = np.zeros(1, dtype=pomdp.state_dtype)[0]
    if name == "rs44":
        s = np.uint32(pomdp.start[0] | (pomdp.start[1] << 4) | (0b1010 << 8))
    elif name == "lightdark":
        s = np.array([(0, 1.5)], dtype=pomdp.state_dtype)[0]
    outcomes = []
    for t in range(steps):
        a_g, N_g, V_g = gp.search(queries)
        rc, a_o, N_o, V_o = op.search(queries)
        assert rc == 0 and a_g == a_o and np.array_equal(N_g, N_o)
        assert np.array_equal(V_g.view(np.uint64), V_o.view(np.uint64))
        sp, o_bits, r = oracle.model_step(pomdp, s, a_g, ut=(rng.random(), 0.0), uo=(rng.random(), rng.random()))
        o = pj.obs_from_bits(pomdp, o_bits)
        rc, out_o = op.update_reinvigorated(a_g, o, seed=900 + t, n_min=n_min)
        if rc == abi.ERR_PARTICLE_DEPLETION:  # posterior without mass (e.g. a terminal belief)
            with pytest.raises(pj.ParticleDepletion):
                gp.update_reinvigorated(a_g, o, 900 + t, n_min)
            break
        assert rc == 0
        out_g = gp.update_reinvigorated(a_g, o, 900 + t, n_min)
        assert out_g == out_o
        assert gp.belief_particles().tobytes() == op.belief_particles().tobytes()
        outcomes.append(out_g)
        if oracle.lib().orc_is_terminal(pomdp.c_struct(), np.array([sp], dtype=pomdp.state_dtype).ctypes.data):
            break
        s = sp
    assert any(o > 0 for o in outcomes), outcomes
    gp.close()


def test_sir_reinvigorator_through_the_api(gpu_lib):
    """POMCPSolver(node_belief_updater=SIRReinvigorator(n)): the reference's depletion scenario
    (tree_queries = 5, 100 steps of Baby) completes through action / update."""
    pomdp = pj.BabyPOMDP(-5.0, -10.0)
    r = pj.SIRReinvigorator(200, rng=4)
    solver = pj.POMCPSolver(c=10.0, rng=2, tree_queries=5, node_belief_updater=r)
    policy = pj.solve(solver, pomdp)
    up = pj.updater(policy)
    b = pj.initialize_belief(up, pj.initial_state_distribution(pomdp))
    rng = np.random.default_rng(1)
    hungry, seen = 0, set()
    for t in range(100):
        a = pj.action(policy, b)
        feed = bool(a)
        hungry_next = 0 if feed else (1 if hungry else int(rng.random() < 0.1))
        cry = int(rng.random() < (0.8 if hungry_next else 0.1))
        b = pj.update(up, b, a, cry)
        seen.add(r.last_outcome)
        n = len(b.particles())
        assert n >= 200 or r.last_outcome == 0
        hungry = hungry_next
    assert seen & {1, 2}
    policy.close()
    # the default updater on the same scenario depletes (the reference's contract)
    policy = pj.solve(pj.POMCPSolver(c=10.0, rng=2, tree_queries=5), pomdp)
    up = pj.updater(policy)
    b = pj.initialize_belief(up, pj.initial_state_distribution(pomdp))
    with pytest.raises(pj.ParticleDepletion):
        for t in range(100):
            a = pj.action(policy, b)
            b = pj.update(up, b, a, int(rng.random() < 0.3))
    policy.close()

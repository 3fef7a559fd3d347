"""Host-side mirror of the POMCP.jl plugin surface over the C ABI.

Same names, argument meaning and error behaviour as the reference
(src/POMCP.jl:20-67 exports): POMCPSolver (src/constructor.jl:8-31), solve /
action (src/solver.jl:16-34), updater / initialize_belief / update
(src/updater.jl:13-47), the NoDecision / AllSamplesTerminal exceptions and
default_action dispatch (src/exceptions.jl:1-26), the particle-depletion
ErrorException (src/particle_filter.jl:94-101) and the SIR particle filter the
reference's tests plug in (test/runtests.jl:57).  All compute goes through
libpomcp_b200.so; nothing here evaluates the algorithm on the CPU.
"""
import ctypes as C

import numpy as np

from . import _abi as abi
from ._lib import lib
from .problems import POMDP, obs_bits, obs_from_bits

__all__ = [
    "POMCPSolver", "POMCPPlanner", "RolloutEstimator", "FORollout", "FOValue", "RandomSolver", "FeedWhenCrying", "DefaultReinvigoratorStub",
    "DeadReinvigorator", "ParticleReinvigorator", "SIRReinvigorator", "ExceptionRethrow", "NoDecision", "AllSamplesTerminal",
    "ParticleDepletion", "PoolExhausted", "POMCPError", "BeliefNode", "RootNode", "ParticleCollection",
    "InitialStateDistribution", "BoolDistribution", "DiscreteBeliefUpdater", "BabyBeliefUpdater", "RootUpdater", "SIRParticleFilter", "solve", "action", "updater", "update",
    "initialize_belief", "initial_state_distribution", "default_action", "root_parallel_action",
    "RolloutSimulator", "simulate", "simulate_episodes", "test_solver", "create_json", "ObsNode", "ActNode", "AbstractActNode",
    "POMCPTreeVisualizer", "AbstractPOMCPSolver", "POMCPDPWSolver", "RandomActionGenerator", "PORollout",
    "PORolloutEstimator", "init_V", "init_N", "sparse_actions", "convert_estimator", "extract_belief", "estimate_value",
]


# ---------------------------------------------------------------- exceptions
class POMCPError(RuntimeError):
    def __init__(self, status, msg=""):
        self.status = status
        super().__init__(f"{abi.STATUS_NAMES.get(status, status)}: {msg}")


class NoDecision(Exception):
    """abstract NoDecision <: Exception (src/exceptions.jl:1-7)."""


class AllSamplesTerminal(NoDecision):
    """src/exceptions.jl:9-18."""

    def __init__(self, belief=None):
        self.belief = belief
        super().__init__("POMCP failed to choose an action because all states sampled from the belief were "
                         "terminal.\n\nTo specify an action for this case, use the default_action solver parameter.")


class ParticleDepletion(RuntimeError):
    """ErrorException of particle_depletion_error() (src/particle_filter.jl:94-101)."""

    def __init__(self):
        super().__init__(
            "POMCP.jl: Particle Depletion! To fix this, you have three options:\n"
            "      1) use more tree_queries (will only work for very small problems)\n"
            "      2) implement a ParticleReinvigorator with reinvigorate!() and handle_unseen_observation()\n"
            "      3) implement a more advanced updater for the agent (POMCP can use any\n"
            "         belief/state distribution that supports rand())")


class PoolExhausted(POMCPError):
    pass


def _check(rc, h=None, belief=None):
    if rc == abi.OK:
        return
    if rc == abi.ERR_ALL_SAMPLES_TERMINAL:
        raise AllSamplesTerminal(belief)
    if rc == abi.ERR_PARTICLE_DEPLETION:
        raise ParticleDepletion()
    msg = lib().pomcp_last_error(h)
    msg = msg.decode() if msg else ""
    if rc == abi.ERR_POOL_EXHAUSTED:
        raise PoolExhausted(rc, msg)
    if rc == abi.ERR_INVALID_ARG:
        raise ValueError(f"pomcp_b200: invalid argument: {msg}")
    if rc == abi.ERR_UNSUPPORTED:
        raise NotImplementedError(f"pomcp_b200: unsupported on device: {msg}")
    if rc == abi.ERR_NAN_VALUE:
        raise AssertionError("!isnan(best_V)")
    raise POMCPError(rc, msg)


class AbstractPOMCPSolver:
    """abstract AbstractPOMCPSolver <: POMDPs.Solver (src/POMCP.jl:73)."""


# ---------------------------------------------------------------- solver pieces
class ExceptionRethrow:
    """src/exceptions.jl:21."""


class RandomSolver:
    """POMDPToolbox.RandomSolver: uniform-random rollout policy (src/constructor.jl:14)."""


class FeedWhenCrying:
    """POMDPModels.FeedWhenCrying: feed iff the previous observation was crying (Baby only)."""


class RolloutEstimator:
    """MCTS.RolloutEstimator(solver_or_policy) (src/rollout.jl:43-46)."""

    def __init__(self, solver=None):
        self.solver = solver if solver is not None else RandomSolver()


class FORollout:
    """FORollout(solver_or_policy) (src/rollout.jl:24-31,86-92): fully observable rollout, the policy is
    given the state.  On device: RandomSolver() (the same trajectories as the PO rollout) and
    FeedWhenCrying() on Baby (feeds iff hungry; test/runtests.jl:25-31)."""

    def __init__(self, solver):
        self.solver = solver


class FOValue:
    """FOValue(policy) (src/rollout.jl:33-39,67-69): the leaf estimate is value(policy, s).  On device the
    policy's value function is a table over the discrete state index (Tiger/Baby: the state; RockSample:
    x | y<<4 | rocks<<8), e.g. the solution of the underlying MDP."""

    def __init__(self, values):
        self.values = np.ascontiguousarray(values, dtype=np.float64)


class ParticleReinvigorator:
    """abstract ParticleReinvigorator (src/particle_filter.jl:37).  Subclasses run on the HOST:
    handle_unseen_observation(old_node, a, o) -> state array (src/particle_filter.jl:60-73),
    reinvigorate(particles, old_node, a, o) -> state array or None (src/particle_filter.jl:46-58)."""

    def reinvigorate(self, particles, old_node, a, o):
        raise NotImplementedError(f"POMCP.jl reinvigorate! not implemented for reinvigorator {type(self)}")

    def handle_unseen_observation(self, old_node, a, o):
        raise NotImplementedError(
            f"POMCP.jl: handle_unseen_observation not implemented for reinvigorator {type(self)}")


class DeadReinvigorator(ParticleReinvigorator):
    """Default: no domain knowledge (src/particle_filter.jl:78,85-92)."""

    def reinvigorate(self, particles, old_node, a, o):
        if len(particles) == 0:
            raise ParticleDepletion()
        return None

    def handle_unseen_observation(self, old_node, a, o):
        raise ParticleDepletion()


class SIRReinvigorator(ParticleReinvigorator):
    """Device-side reinvigorator: both hooks draw from the SIR posterior of the old root belief
    (propagate with generate_s, weight with the observation pdf, systematic resample) so that the new
    belief holds at least `n` particles; an unseen observation rebuilds the belief from `n` of them
    (src/particle_filter.jl:37-73).  Runs inside pomcp_update_reinvigorated."""

    def __init__(self, n, rng=0):
        self.n, self.seed, self._step = int(n), int(rng), 0
        self.last_outcome = None  # 0 plain re-root, 1 topped up, 2 rebuilt from the posterior

    def next_seed(self):
        self._step += 1
        return (self.seed * 0x9E3779B97F4A7C15 + self._step) & (2 ** 64 - 1)


class DiscreteBeliefUpdater:
    """The problem's own exact belief updater (`updater(problem)`: POMDPModels.BabyBeliefUpdater, the
    DiscreteUpdater of Tiger) used as node_belief_updater (notebooks/Sanity_Checks.ipynb:47,70;
    test/Belief_and_Particle_Filter_Options.jl:63-77): every tree node then carries the exact belief
    (src/solver.jl:138) and update() never depletes (src/updater.jl:30-39).  On device for the two-state
    problems (Tiger, Baby), where the belief is the single number P(s = 1)."""

    def __init__(self, pomdp=None):
        self.pomdp = pomdp


BabyBeliefUpdater = DiscreteBeliefUpdater


class BoolDistribution:
    """POMDPModels.BoolDistribution(p): P(true) = p (hungry / tiger behind the left door)."""

    def __init__(self, p):
        self.p = float(p)

    def __repr__(self):
        return f"BoolDistribution({self.p})"


class DefaultReinvigoratorStub:
    """src/particle_filter.jl:83; replaced by DeadReinvigorator in solve() (src/solver.jl:6-10)."""


class POMCPSolver(AbstractPOMCPSolver):
    """Keyword constructor with the reference defaults (src/constructor.jl:8-18; fields documented
    at src/POMCP.jl:75-136).  `rng` is an integer seed (Philox key) instead of an AbstractRNG.
    Device-only keywords: search_mode ("batched"|"exact"), inflight_min, inflight_max, inflight_ramp_div, rollouts_per_leaf,
    max_nodes, max_log, device, rank, n_trees."""

    def __init__(self, eps=0.01, max_depth=2 ** 63 - 1, c=1.0, tree_queries=100, rng=0,
                 node_belief_updater=None, estimate_value=None, init_V=0.0, init_N=0, num_sparse_actions=0,
                 default_action=None, search_mode="batched", inflight_min=0, inflight_max=0, inflight_ramp_div=0, rollouts_per_leaf=0,
                 max_nodes=0,
                 max_log=0, device=0, rank=0, dpw_observation=None, dpw_action=None, dpw_solver=False, n_trees=1):
        self.eps, self.max_depth, self.c, self.tree_queries = eps, max_depth, c, tree_queries
        self.rng = rng
        self.node_belief_updater = node_belief_updater if node_belief_updater is not None \
            else DefaultReinvigoratorStub()
        self.estimate_value = estimate_value if estimate_value is not None else RolloutEstimator(RandomSolver())
        self.init_V, self.init_N, self.num_sparse_actions = init_V, init_N, num_sparse_actions
        self.default_action = default_action if default_action is not None else ExceptionRethrow()
        self.search_mode = search_mode
        self.inflight_min, self.inflight_max, self.inflight_ramp_div = inflight_min, inflight_max, inflight_ramp_div
        self.rollouts_per_leaf = rollouts_per_leaf
        self.max_nodes, self.max_log, self.device, self.rank = max_nodes, max_log, device, rank
        self.dpw_observation = dpw_observation  # (k_observation, alpha_observation) of POMCPDPWSolver, or None
        self.dpw_action = dpw_action  # (k_action, alpha_action) of POMCPDPWSolver with enable_action_pw, or None
        self.dpw_solver = dpw_solver  # simulate() of POMCPDPWSolver (its leaf step budget and untried-action rule)
        self.n_trees = n_trees  # root parallelism inside the GPU: trees of tree_queries simulations each, merged

    def c_params(self, belief_mode):
        sp = abi.SolverParams()
        lib().pomcp_default_solver_params(C.byref(sp))
        sp.eps, sp.max_depth, sp.c, sp.tree_queries = self.eps, min(int(self.max_depth), 2 ** 63 - 1), self.c, \
            self.tree_queries
        sp.seed = int(self.rng) & (2 ** 64 - 1)
        if callable(self.init_V) or callable(self.init_N):
            raise NotImplementedError("init_V/init_N callbacks cannot run on device (src/tree.jl:35,44)")
        sp.init_V, sp.init_N = float(self.init_V), int(self.init_N)
        sp.num_sparse_actions = int(self.num_sparse_actions)
        ev = self.estimate_value
        if isinstance(ev, RolloutEstimator) and isinstance(ev.solver, RandomSolver):  # includes PORollout
            sp.estimator = abi.EST_RANDOM_ROLLOUT
        elif isinstance(ev, RolloutEstimator) and isinstance(ev.solver, FeedWhenCrying):
            sp.estimator = abi.EST_FWC_ROLLOUT  # test/runtests.jl:13-19, notebooks/Display_Tree.ipynb:31-39
        elif isinstance(ev, FORollout) and isinstance(ev.solver, RandomSolver):
            sp.estimator = abi.EST_RANDOM_ROLLOUT  # a random policy ignores what it is shown
        elif isinstance(ev, FORollout) and isinstance(ev.solver, FeedWhenCrying):
            sp.estimator = abi.EST_FO_FWC_ROLLOUT  # test/runtests.jl:25-31
        elif isinstance(ev, FOValue):
            sp.estimator = abi.EST_STATE_VALUE
        elif isinstance(ev, (int, float)):
            sp.estimator, sp.estimator_value = abi.EST_CONSTANT, float(ev)  # src/rollout.jl:10
        else:
            raise NotImplementedError(
                "estimate_value: RolloutEstimator / FORollout over RandomSolver() or FeedWhenCrying(), FOValue(table) "
                "and numeric estimates run on device (function estimators and arbitrary policies are host "
                "callbacks, src/rollout.jl:9,24-39)")
        sp.belief_mode = belief_mode
        sp.search_mode = {"exact": abi.MODE_EXACT, "batched": abi.MODE_BATCHED}[self.search_mode]
        sp.max_nodes, sp.max_log = int(self.max_nodes), int(self.max_log)
        sp.inflight_min, sp.inflight_max = self.inflight_min, self.inflight_max
        sp.inflight_ramp_div, sp.rollouts_per_leaf = self.inflight_ramp_div, self.rollouts_per_leaf
        sp.rank = self.rank
        if self.dpw_observation is not None:
            sp.dpw_observation = 1
            sp.k_observation, sp.alpha_observation = float(self.dpw_observation[0]), float(self.dpw_observation[1])
        sp.dpw_solver = 1 if self.dpw_solver else 0
        sp.n_trees = int(self.n_trees)
        if self.dpw_action is not None:
            sp.dpw_action = 1
            sp.k_action, sp.alpha_action = float(self.dpw_action[0]), float(self.dpw_action[1])
        return sp


# ---------------------------------------------------------------- beliefs
class ParticleCollection:
    """ParticleFilters.ParticleCollection: unweighted particle belief (src/particle_filter.jl:1-18)."""

    def __init__(self, particles):
        self.particles = np.ascontiguousarray(particles)

    def __len__(self):
        return len(self.particles)


class InitialStateDistribution:
    """initial_state_distribution(pomdp): sampled on device by the model's own rule."""

    def __init__(self, pomdp):
        self.pomdp = pomdp


def initial_state_distribution(pomdp):
    return InitialStateDistribution(pomdp)


class AbstractActNode:
    """abstract AbstractActNode (src/tree.jl:1-5)."""


class ActNode(AbstractActNode):
    """View of an action node of the device tree (src/tree.jl:7-12): label, N, V, children[o] -> ObsNode."""

    def __init__(self, parent, a, N, V):
        self._parent, self.label, self.N, self.V = parent, int(a), int(N), float(V)

    @property
    def children(self):
        return _ObsChildren(self._parent, self.label)

    def __iter__(self):  # the (N, V) pair earlier versions of this mirror returned
        return iter((self.N, self.V))

    def __repr__(self):
        return f"ActNode(label={self.label}, N={self.N}, V={self.V:.6g})"


class _ObsChildren:
    """children::Dict{O,ObsNode} of an ActNode, looked up on the device (pomcp_get_node_stats)."""

    def __init__(self, parent, a):
        self._parent, self._a = parent, a

    def _node(self, o):
        return ObsNode(self._parent.planner, self._parent.generation, history=self._parent._history + [(self._a, o)],
                       label=o)

    def __contains__(self, o):
        return self._node(o)._exists()

    def __getitem__(self, o):
        n = self._node(o)
        if not n._exists():
            raise KeyError(o)
        return n

    def get(self, o, default=None):
        n = self._node(o)
        return n if n._exists() else default

    def keys(self):
        pomdp = self._parent.planner.problem
        if pomdp.continuous_obs:
            raise NotImplementedError("children of continuous observations cannot be enumerated through the ABI")
        return [o for o in range(pomdp.n_observations) if o in self]

    def __iter__(self):
        return iter(self.keys())

    def __len__(self):
        return len(self.keys())

    def values(self):
        return [self[o] for o in self.keys()]

    def items(self):
        return [(o, self[o]) for o in self.keys()]


class BeliefNode:
    """Handle to a belief node of the device tree (src/tree.jl:14-26): the planner's current root (it is
    both the belief and the search root, src/updater.jl:20-26) or, through `children`, any node below it.
    `N`, `children[a] -> ActNode`, `children[a].children[o] -> ObsNode` read the device tree through
    pomcp_get_node_stats, so the tree can be walked like the reference's."""

    def __init__(self, planner, generation, B=None, history=None, label=None):
        self.planner, self.generation, self.B = planner, generation, B
        self._history = list(history or [])
        self.label = label

    def _live(self):
        if self.generation != self.planner._generation:
            raise ValueError("stale BeliefNode: the planner's tree has been re-rooted since")

    def _stats(self):
        self._live()
        return self.planner.node_stats(self._history)

    def _exists(self):
        return self._stats()[0] >= 0

    @property
    def N(self):
        return self._stats()[0]

    @property
    def children(self):
        """{action: ActNode} (all actions are expanded together, src/solver.jl:87-94)."""
        _, _, N, V = self._stats()
        return {a: ActNode(self, a, N[a], V[a]) for a in range(len(N))}

    def particles(self):
        self._live()
        if self._history:
            raise NotImplementedError("only the root's particle collection is materialised (pomcp_get_belief_particles)")
        return self.planner.belief_particles()


class RootNode(BeliefNode):
    pass


class ObsNode(BeliefNode):
    pass


class POMCPTreeVisualizer:
    """POMCPTreeVisualizer(node) (src/visualization.jl:4-6); create_json(visualizer) exports the tree."""

    def __init__(self, node):
        self.node = node


# ---------------------------------------------------------------- planner
class POMCPPlanner:
    """POMCPPlanner (src/POMCP.jl:260-268): owns the opaque device handle."""

    def __init__(self, solver, pomdp, belief_mode=abi.BELIEF_UNWEIGHTED):
        if not isinstance(pomdp, POMDP):
            raise NotImplementedError("only the built-in problem models run on device")
        self.solver, self.problem = solver, pomdp
        nbu = solver.node_belief_updater
        if isinstance(nbu, DefaultReinvigoratorStub):
            nbu = DeadReinvigorator()  # src/solver.jl:6-10
        if isinstance(nbu, DiscreteBeliefUpdater):
            belief_mode = abi.BELIEF_EXACT2  # the generic branch of src/solver.jl:138 / src/updater.jl:30-39
        elif not isinstance(nbu, ParticleReinvigorator):
            raise NotImplementedError("node_belief_updater: ParticleReinvigorators and the exact updater of the two-state "
                                      "problems (DiscreteBeliefUpdater) are supported on device")
        self.node_belief_updater = nbu
        self.belief_mode = belief_mode
        self._p = pomdp.c_struct()
        self._sp = solver.c_params(belief_mode)
        self.nA = pomdp.n_actions
        h = C.c_void_p()
        rc = lib().pomcp_create(C.byref(self._p), C.byref(self._sp), int(solver.device), C.byref(h))
        if rc != abi.OK:
            _check(rc, None)
        self._h = h
        if isinstance(solver.estimate_value, FOValue):
            v = solver.estimate_value.values
            _check(lib().pomcp_set_state_values(self._h, v.ctypes.data_as(C.c_void_p), len(v)), self._h)
        self._generation = 0
        self._tree_ref = None  # src/POMCP.jl:266-267

    def close(self):
        if getattr(self, "_h", None):
            lib().pomcp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- thin wrappers over the C ABI
    def set_belief_particles(self, states):
        states = np.ascontiguousarray(states, dtype=self.problem.state_dtype)
        _check(lib().pomcp_set_belief_particles(self._h, states.ctypes.data_as(C.c_void_p), len(states)), self._h)
        self._generation += 1

    def set_belief_initial(self):
        _check(lib().pomcp_set_belief_initial(self._h), self._h)
        self._generation += 1

    def set_belief_exact2(self, p1):
        _check(lib().pomcp_set_belief_exact2(self._h, float(p1)), self._h)
        self._generation += 1

    def belief_exact2(self):
        p = C.c_double(0.0)
        _check(lib().pomcp_get_belief_exact2(self._h, C.byref(p)), self._h)
        return p.value

    def update_exact2(self, a, o):
        _check(lib().pomcp_update_exact2(self._h, int(a), obs_bits(o)), self._h)
        self._generation += 1

    def reset_tree(self):
        """A fresh RootNode over the belief already on the device (make_node, src/solver.jl:278-281)."""
        _check(lib().pomcp_reset_tree(self._h), self._h)
        self._generation += 1

    def belief_particles(self, out=None):
        """The belief's particles as a host array; `out` (same dtype, at least as long — e.g. a pinned buffer
        the caller re-uses from step to step) receives them instead of a fresh array."""
        n = C.c_int64(0)
        _check(lib().pomcp_get_belief_particles(self._h, None, 0, C.byref(n)), self._h)
        if out is not None:
            if out.dtype != self.problem.state_dtype or out.ndim != 1 or len(out) < n.value or not out.flags.c_contiguous:
                raise ValueError("out must be a contiguous 1-d array of the problem's state dtype, long enough")
            out = out[:n.value]
        else:
            out = np.zeros(n.value, dtype=self.problem.state_dtype)
        if n.value:
            _check(lib().pomcp_get_belief_particles(self._h, out.ctypes.data_as(C.c_void_p), n.value, C.byref(n)),
                   self._h)
        return out

    def search(self, tree_queries=-1, belief=None):
        a = C.c_int32(-1)
        N = np.zeros(self.nA, dtype=np.int64)
        V = np.zeros(self.nA, dtype=np.float64)
        rc = lib().pomcp_search(self._h, tree_queries, C.byref(a), N.ctypes.data_as(C.c_void_p),
                                V.ctypes.data_as(C.c_void_p), self.nA)
        _check(rc, self._h, belief)
        return a.value, N, V

    def search_replay(self, tree_queries, uniforms, returns):
        a = C.c_int32(-1)
        N = np.zeros(self.nA, dtype=np.int64)
        V = np.zeros(self.nA, dtype=np.float64)
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        returns = np.ascontiguousarray(returns, dtype=np.float64)
        uu, ru = C.c_int64(0), C.c_int64(0)
        rc = lib().pomcp_search_replay(self._h, tree_queries, uniforms.ctypes.data_as(C.c_void_p), len(uniforms),
                                       returns.ctypes.data_as(C.c_void_p), len(returns), C.byref(a),
                                       N.ctypes.data_as(C.c_void_p), V.ctypes.data_as(C.c_void_p), self.nA,
                                       C.byref(uu), C.byref(ru))
        _check(rc, self._h)
        return a.value, N, V, uu.value, ru.value

    def search_rootparallel(self, tree_queries=-1):
        a = C.c_int32(-1)
        N = np.zeros(self.nA, dtype=np.int64)
        V = np.zeros(self.nA, dtype=np.float64)
        rc = lib().pomcp_search_rootparallel(self._h, tree_queries, C.byref(a), N.ctypes.data_as(C.c_void_p),
                                             V.ctypes.data_as(C.c_void_p), self.nA)
        _check(rc, self._h)
        return a.value, N, V

    def update_unweighted(self, a, o):
        rc = lib().pomcp_update_unweighted(self._h, int(a), obs_bits(o))
        _check(rc, self._h)
        self._generation += 1

    def update_reinvigorated(self, a, o, seed, n_min):
        out = C.c_int32(-1)
        rc = lib().pomcp_update_reinvigorated(self._h, int(a), obs_bits(o), int(seed), int(n_min), C.byref(out))
        _check(rc, self._h)
        self._generation += 1
        return out.value

    def pf_update_sir(self, a, o, seed, n_out):
        _check(lib().pomcp_pf_update_sir(self._h, int(a), obs_bits(o), int(seed), int(n_out)), self._h)
        self._generation += 1

    def root_stats(self):
        rn = C.c_int64(0)
        N = np.zeros(self.nA, dtype=np.int64)
        V = np.zeros(self.nA, dtype=np.float64)
        _check(lib().pomcp_get_root_stats(self._h, C.byref(rn), N.ctypes.data_as(C.c_void_p),
                                          V.ctypes.data_as(C.c_void_p), self.nA), self._h)
        return rn.value, N, V

    def tree_result(self, tree):
        """(action, root N, N[a], V[a]) of tree `tree` of the last search, before the merge over trees / ranks."""
        a, rn = C.c_int32(-1), C.c_int64(0)
        N = np.zeros(self.nA, dtype=np.int64)
        V = np.zeros(self.nA, dtype=np.float64)
        _check(lib().pomcp_get_tree_result(self._h, int(tree), C.byref(a), C.byref(rn), N.ctypes.data_as(C.c_void_p),
                                           V.ctypes.data_as(C.c_void_p), self.nA), self._h)
        return a.value, rn.value, N, V

    def node_stats(self, history):
        acts = np.array([h[0] for h in history], dtype=np.int32)
        obs = np.array([obs_bits(h[1]) for h in history], dtype=np.uint64)
        nN, npart = C.c_int64(0), C.c_int64(0)
        N = np.zeros(self.nA, dtype=np.int64)
        V = np.zeros(self.nA, dtype=np.float64)
        _check(lib().pomcp_get_node_stats(self._h, acts.ctypes.data_as(C.c_void_p), obs.ctypes.data_as(C.c_void_p),
                                          len(history), C.byref(nN), C.byref(npart), N.ctypes.data_as(C.c_void_p),
                                          V.ctypes.data_as(C.c_void_p), self.nA), self._h)
        return nN.value, npart.value, N, V

    def root_node(self):
        """index of the current root in the dump_tree arrays"""
        n = C.c_int64(0)
        _check(lib().pomcp_root_node(self._h, C.byref(n)), self._h)
        return n.value

    def dump_tree(self):
        n = C.c_int64(0)
        _check(lib().pomcp_dump_tree(self._h, 0, C.byref(n), None, None, None, None), self._h)
        nn = n.value
        nO = getattr(self.problem, "n_observations", 0)
        node_N = np.zeros(nn, dtype=np.int64)
        act_N = np.zeros((nn, self.nA), dtype=np.int64)
        act_V = np.zeros((nn, self.nA), dtype=np.float64)
        child = np.full((nn, self.nA, max(nO, 1)), -1, dtype=np.int32) if nO else None
        _check(lib().pomcp_dump_tree(self._h, nn, C.byref(n), node_N.ctypes.data_as(C.c_void_p),
                                     act_N.ctypes.data_as(C.c_void_p), act_V.ctypes.data_as(C.c_void_p),
                                     None if child is None else child.ctypes.data_as(C.c_void_p)), self._h)
        return node_N, act_N, act_V, child

    def generate_sor(self, state, a, seed, step):
        """generate_sor(pomdp, s, a, rng) on the device model -> (sp, o, r, terminal)."""
        s = np.ascontiguousarray(np.array([state], dtype=self.problem.state_dtype))
        sp = np.zeros(1, dtype=self.problem.state_dtype)
        o, r, t = C.c_uint64(0), C.c_double(0.0), C.c_int32(0)
        _check(lib().pomcp_generate_sor(self._h, s.ctypes.data_as(C.c_void_p), int(a), int(seed), int(step),
                                        sp.ctypes.data_as(C.c_void_p), C.byref(o), C.byref(r), C.byref(t)), self._h)
        return sp[0], obs_from_bits(self.problem, o.value), r.value, bool(t.value)

    def run_episodes(self, n_episodes, max_steps, env_seed=0, initial_state=-1, sim_eps=0.0, sir_particles=0,
                     initial_p1=0.0):
        """n_episodes runs of the outer simulation loop on the device, one tree per episode, no host round trip
        inside an episode (pomcp_run_episodes).  Returns (discounted rewards, steps taken, status per episode)."""
        ep = abi.EpisodeParams(int(max_steps), int(initial_state), float(sim_eps), int(env_seed) & (2 ** 64 - 1),
                               int(sir_particles), float(initial_p1))
        r = np.zeros(n_episodes, dtype=np.float64)
        st = np.zeros(n_episodes, dtype=np.int32)
        status = np.zeros(n_episodes, dtype=np.int32)
        _check(lib().pomcp_run_episodes(self._h, C.byref(ep), int(n_episodes), r.ctypes.data_as(C.c_void_p),
                                        st.ctypes.data_as(C.c_void_p), status.ctypes.data_as(C.c_void_p)), self._h)
        self._generation += 1
        return r, st, status

    def initial_state(self, seed):
        s = np.zeros(1, dtype=self.problem.state_dtype)
        _check(lib().pomcp_initial_state(self._h, int(seed), s.ctypes.data_as(C.c_void_p)), self._h)
        return s[0]

    def counters(self):
        c = abi.Counters()
        _check(lib().pomcp_get_counters(self._h, C.byref(c)), self._h)
        return c.as_dict()

    def worker_cycles(self):
        out = np.zeros(8, dtype=np.uint64)
        _check(lib().pomcp_worker_cycles(self._h, out.ctypes.data_as(C.c_void_p)), self._h)
        return out

    def reset_counters(self):
        _check(lib().pomcp_reset_counters(self._h), self._h)

    def phase_ms(self, phase):
        ms, n = C.c_double(0.0), C.c_int64(0)
        _check(lib().pomcp_last_phase_ms(self._h, phase, C.byref(ms), C.byref(n)), self._h)
        return ms.value, n.value

    def event_record(self, slot):
        _check(lib().pomcp_event_record(self._h, slot), self._h)

    def event_elapsed_ms(self, s0, s1):
        ms = C.c_double(0.0)
        _check(lib().pomcp_event_elapsed_ms(self._h, s0, s1, C.byref(ms)), self._h)
        return ms.value

    def search_async(self, tree_queries=-1):
        _check(lib().pomcp_search_async(self._h, tree_queries), self._h)

    def sync(self):
        rc = lib().pomcp_sync(self._h)
        _check(rc, self._h)

    def flush_l2(self):
        _check(lib().pomcp_flush_l2(self._h), self._h)

    def rollout_batch(self, states, steps, seed):
        states = np.ascontiguousarray(states, dtype=self.problem.state_dtype)
        ret = np.zeros(len(states), dtype=np.float64)
        ns = np.zeros(len(states), dtype=np.int32)
        _check(lib().pomcp_rollout_batch(self._h, states.ctypes.data_as(C.c_void_p), len(states), int(steps),
                                         int(seed), ret.ctypes.data_as(C.c_void_p), ns.ctypes.data_as(C.c_void_p)),
               self._h)
        return ret, ns

    def pf_weights(self, states, a, o, seed):
        states = np.ascontiguousarray(states, dtype=self.problem.state_dtype)
        w = np.zeros(len(states), dtype=np.float64)
        nxt = np.zeros(len(states), dtype=self.problem.state_dtype)
        _check(lib().pomcp_pf_weights(self._h, states.ctypes.data_as(C.c_void_p), len(states), int(a), obs_bits(o),
                                      int(seed), w.ctypes.data_as(C.c_void_p), nxt.ctypes.data_as(C.c_void_p)),
               self._h)
        return w, nxt

    def resample_systematic(self, weights, r, n_out, mode=0):
        weights = np.ascontiguousarray(weights, dtype=np.float64)
        idx = np.zeros(n_out, dtype=np.int64)
        rc = lib().pomcp_resample_systematic(self._h, weights.ctypes.data_as(C.c_void_p), len(weights), float(r),
                                             int(n_out), int(mode), idx.ctypes.data_as(C.c_void_p))
        _check(rc, self._h)
        return idx

    def comm_init(self, unique_id: bytes, rank, nranks):
        buf = C.create_string_buffer(unique_id, 128)
        _check(lib().pomcp_comm_init(self._h, buf, rank, nranks), self._h)


def nccl_unique_id() -> bytes:
    buf = C.create_string_buffer(128)
    _check(lib().pomcp_nccl_unique_id(buf))
    return buf.raw


# ---------------------------------------------------------------- POMDPs.jl verbs
def solve(solver, pomdp, belief_mode=abi.BELIEF_UNWEIGHTED):
    """solve(solver, pomdp) -> POMCPPlanner (src/solver.jl:30-32)."""
    if isinstance(solver, POMCPDPWSolver):
        solver = solver.as_pomcp(pomdp)
        if solver.dpw_observation is not None:
            belief_mode = abi.BELIEF_EXTERNAL  # continuous observations never re-root
    if not isinstance(solver, POMCPSolver):
        raise NotImplementedError(f"{type(solver).__name__} has no device path")
    return POMCPPlanner(solver, pomdp, belief_mode)


def default_action(spec, belief, ex):
    """src/exceptions.jl:23-26."""
    if isinstance(spec, ExceptionRethrow):
        raise ex
    if isinstance(spec, POMCPPlanner):
        return action(spec, belief)
    if callable(spec):
        return spec(belief, ex)
    return spec


def _make_node(planner, belief):
    """make_node (src/solver.jl:276-281): a BeliefNode is searched in place, anything else gets
    a fresh RootNode."""
    if isinstance(belief, BeliefNode):
        if belief.planner is not planner:
            raise ValueError("BeliefNode belongs to another planner")
        belief._live()
        return belief
    if isinstance(belief, ParticleCollection):
        planner.set_belief_particles(belief.particles)
    elif isinstance(belief, BoolDistribution):
        planner.set_belief_exact2(belief.p)
    elif isinstance(belief, InitialStateDistribution) and planner.belief_mode == abi.BELIEF_EXACT2:
        planner.set_belief_exact2(belief.pomdp.initial_p1)  # the distribution itself, not samples of it
    elif isinstance(belief, InitialStateDistribution):
        planner.set_belief_initial()
    elif isinstance(belief, np.ndarray):
        planner.set_belief_particles(belief)
    else:
        raise NotImplementedError("belief must be a BeliefNode, ParticleCollection or initial_state_distribution "
                                  "(arbitrary rand(rng, b) callbacks cannot run on device)")
    return RootNode(planner, planner._generation, belief)


def action(policy, belief):
    """action(policy, belief) (src/solver.jl:16-23)."""
    try:
        node = _make_node(policy, belief)
        policy._tree_ref = node
        a, _, _ = policy.search(policy.solver.tree_queries, belief)
        return a
    except (NoDecision, ParticleDepletion, AssertionError, PoolExhausted) as ex:
        return default_action(policy.solver.default_action, belief, ex)


class RootUpdater:
    """RootUpdater (src/updater.jl:8-10)."""

    def __init__(self, planner):
        self.planner = planner
        self.node_belief_updater = planner.node_belief_updater


def updater(policy):
    """updater(policy::POMCPPlanner) = RootUpdater (src/updater.jl:41); updater(problem) = the problem's exact
    belief updater (the node_belief_updater of notebooks/Sanity_Checks.ipynb:47)."""
    if isinstance(policy, POMDP):
        return DiscreteBeliefUpdater(policy)
    return RootUpdater(policy)


def initialize_belief(up, b):
    """initialize_belief(up::RootUpdater, b) = RootNode(0, b, {}) (src/updater.jl:45-47)."""
    if isinstance(up, SIRParticleFilter):
        return up.initialize_belief(b)
    if isinstance(b, BeliefNode):
        return b
    return _make_node(up.planner, b)


def update(up, b_old, a, o, out=None):
    """update(updater, b_old, a, o) (src/updater.jl:13-27 for RootUpdater; SIR filter otherwise).
    `out`: SIR filter only, host buffer that receives the new particle set (see belief_particles)."""
    if isinstance(up, SIRParticleFilter):
        return up.update(b_old, a, o, out=out)
    pl = up.planner
    if not isinstance(b_old, BeliefNode):
        raise TypeError("RootUpdater.update needs the BeliefNode returned by initialize_belief/update")
    b_old._live()
    if isinstance(up.node_belief_updater, DiscreteBeliefUpdater):  # src/updater.jl:30-39
        pl.update_exact2(a, o)
        node = BeliefNode(pl, pl._generation)
        node.B = BoolDistribution(pl.belief_exact2())
        return node
    if isinstance(up.node_belief_updater, SIRReinvigorator):  # both hooks run on the device
        r = up.node_belief_updater
        r.last_outcome = pl.update_reinvigorated(a, o, r.next_seed(), r.n)
        return BeliefNode(pl, pl._generation)
    try:
        pl.update_unweighted(a, o)
    except ParticleDepletion:
        r = up.node_belief_updater
        if isinstance(r, DeadReinvigorator):
            raise
        new_particles = r.handle_unseen_observation(b_old, a, o)  # src/updater.jl:14-18 (host hook)
        pl.set_belief_particles(new_particles)
        return BeliefNode(pl, pl._generation)
    node = BeliefNode(pl, pl._generation)
    r = up.node_belief_updater
    if not isinstance(r, DeadReinvigorator):
        extra = r.reinvigorate(node.particles(), b_old, a, o)  # src/updater.jl:21-23 (host hook)
        if extra is not None and len(extra):
            raise NotImplementedError("adding particles to a re-rooted device node is not supported yet")
    return node


class SIRParticleFilter:
    """ParticleFilters.SIRParticleFilter(model, n) (test/runtests.jl:57): propagate, weight,
    systematic resample on the device.  Beliefs are ParticleCollections, so the planner builds a
    fresh root every step (src/solver.jl:278-281)."""

    def __init__(self, planner, n, rng=0):
        self.planner, self.n, self.seed, self._step = planner, int(n), int(rng), 0

    def initialize_belief(self, b):
        if isinstance(b, InitialStateDistribution):
            raise NotImplementedError("draw the initial particles on the host and pass a ParticleCollection")
        return b if isinstance(b, ParticleCollection) else ParticleCollection(b)

    def update(self, b, a, o, on_device=False, out=None):
        pl = self.planner
        if not on_device:
            pl.set_belief_particles(b.particles)
        self._step += 1
        pl.pf_update_sir(a, o, (self.seed * 0x9E3779B97F4A7C15 + self._step) & (2 ** 64 - 1), self.n)
        return ParticleCollection(pl.belief_particles(out))


def root_parallel_action(policy):
    """Root-parallel decision: every rank searches its own tree, root N / N*V are merged by one
    NCCL all-reduce inside pomcp_search_rootparallel (needs comm_init on every rank)."""
    return policy.search_rootparallel(policy.solver.tree_queries)[0]


# ---------------------------------------------------------------- outer simulation loop (L4)
class RolloutSimulator:
    """POMDPToolbox.RolloutSimulator as the reference's tests and notebooks use it (test_solver,
    notebooks/Basic_Usage.ipynb:88-103; loop body SURVEY.md App. B.1): runs one episode and returns the
    discounted reward.  The environment steps with pomcp_generate_sor, i.e. the same model code as the
    planner, on its own Philox stream `rng`."""

    def __init__(self, rng=0, initial_state=None, eps=None, max_steps=None):
        self.rng, self.initial_state, self.eps, self.max_steps = int(rng), initial_state, eps, max_steps
        self.history = []  # (s, a, o, r) of the last episode


def simulate(sim, pomdp, policy, up=None, initial_belief=None):
    """simulate(sim::RolloutSimulator, pomdp, policy, updater, initial_belief) -> discounted reward."""
    if policy.problem is not pomdp:
        raise ValueError("the planner was solved for another problem instance")
    up = up if up is not None else updater(policy)
    dist = initial_belief if initial_belief is not None else initial_state_distribution(pomdp)
    s = sim.initial_state if sim.initial_state is not None else policy.initial_state(sim.rng)
    b = initialize_belief(up, dist)
    eps = 0.0 if sim.eps is None else sim.eps
    max_steps = 2 ** 62 if sim.max_steps is None else sim.max_steps
    disc, total, step, terminal = 1.0, 0.0, 1, False
    sim.history = []
    while disc > eps and not terminal and step <= max_steps:
        a = action(policy, b)
        sp, o, r, terminal = policy.generate_sor(s, a, sim.rng, step)
        sim.history.append((s, a, o, r))
        total += disc * r
        s = sp
        if not terminal and step < max_steps:
            b = update(up, b, a, o)
        disc *= pomdp.discount
        step += 1
    return total


def simulate_episodes(sim, pomdp, policy, n_episodes, up=None, initial_belief=None):
    """n_episodes independent runs of simulate(sim, pomdp, policy, updater, initial_belief) - what est_reward of
    notebooks/Sanity_Checks.ipynb:47-77 spreads over worker processes - in ONE device call: one tree per episode
    (the planner needs n_trees >= n_episodes), no host round trip inside an episode.  Episode e draws its
    environment from Philox(sim.rng + e).  Returns (discounted rewards, steps, status codes)."""
    import math
    if policy.problem is not pomdp:
        raise ValueError("the planner was solved for another problem instance")
    eps = 0.0 if sim.eps is None else float(sim.eps)
    if sim.max_steps is not None:
        max_steps = int(sim.max_steps)
    elif eps > 0.0:
        max_steps = int(math.ceil(math.log(eps) / math.log(pomdp.discount)))  # disc > eps fails after that many steps
    else:
        raise ValueError("RolloutSimulator needs max_steps or eps for a device-resident episode loop")
    kw = {}
    if isinstance(up, SIRParticleFilter):
        if policy.belief_mode != abi.BELIEF_EXTERNAL:
            raise ValueError("a SIR particle filter needs a planner solved with belief_mode = EXTERNAL")
        kw["sir_particles"] = up.n
    elif policy.belief_mode == abi.BELIEF_EXACT2:
        b0 = initial_belief if initial_belief is not None else initial_state_distribution(pomdp)
        kw["initial_p1"] = b0.p if isinstance(b0, BoolDistribution) else pomdp.initial_p1
    s0 = -1 if sim.initial_state is None else int(sim.initial_state)
    return policy.run_episodes(n_episodes, max_steps, env_seed=sim.rng, initial_state=s0, sim_eps=eps, **kw)


def test_solver(solver, pomdp, max_steps=10, updater=None, rng=0):
    """POMDPToolbox.test_solver: solve, run a short episode, return its reward."""
    policy = solve(solver, pomdp)
    try:
        up = updater(policy) if callable(updater) else updater
        return simulate(RolloutSimulator(rng=rng, max_steps=max_steps), pomdp, policy, up)
    finally:
        policy.close()


test_solver.__test__ = False  # not a pytest test


def create_json(policy):
    if isinstance(policy, POMCPTreeVisualizer):
        policy = policy.node.planner
    elif isinstance(policy, BeliefNode):
        policy = policy.planner
    return _create_json(policy)


def _create_json(policy):
    """create_json(POMCPTreeVisualizer(node)) (src/visualization.jl:12-72): the tree under the planner's root
    as the {id: {id, type, children_ids, tag, tt_tag, N[, V]}} dictionary MCTS.jl's D3 widget reads,
    serialised to JSON; returns (json, root_id) like the reference."""
    import json
    node_N, act_N, act_V, child = policy.dump_tree()
    if child is None:
        raise NotImplementedError("tree export needs discrete observations")
    nd = {}

    def push_belief(n, label, parent):
        i = len(nd) + 1
        if parent:
            nd[parent]["children_ids"].append(i)
        nd[i] = {"id": i, "type": "belief", "children_ids": [], "tag": str(label), "tt_tag": str(label),
                 "N": int(node_N[n])}
        if act_N[n].sum() == 0 and node_N[n] == 0:
            return
        for a in range(act_N.shape[1]):
            j = len(nd) + 1
            nd[i]["children_ids"].append(j)
            nd[j] = {"id": j, "type": "action", "children_ids": [], "tag": str(a), "tt_tag": str(a),
                     "N": int(act_N[n, a]), "V": float(act_V[n, a])}
            for o in range(child.shape[2]):
                if child[n, a, o] >= 0:
                    push_belief(int(child[n, a, o]), o, j)

    import sys
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 10000))
    push_belief(policy.root_node(), "root", 0)
    return json.dumps(nd), 1


# ---------------------------------------------------------------- POMCPDPWSolver and the small exported helpers the tests call
class POMCPDPWSolver(AbstractPOMCPSolver):
    """POMCPDPWSolver (src/POMCP.jl:154-253, src/constructor.jl:40-73, src/solver.jl:161-274): double progressive
    widening, run on the device through POMCPSolver's kernels with the widening switches of the parameter block:
    action widening (`enable_action_pw`, next_action = RandomActionGenerator) on every problem, observation
    widening on continuous observations (LightDark1D, test/runtests.jl:57; needs an external belief updater, as
    in the reference's test).  On a discrete observation space of at most `k_observation` elements the test
    `length(children) <= k_observation * N^alpha_observation` (src/solver.jl:223) is always true, so no
    observation widening is needed there (test/runtests.jl:44-54)."""

    def __init__(self, eps=0.01, max_depth=2 ** 63 - 1, c=1.0, tree_queries=100, rng=0, node_belief_updater=None,
                 estimate_value=None, enable_action_pw=True, alpha_observation=0.5, k_observation=10.0,
                 alpha_action=0.5, k_action=10.0, init_V=0.0, init_N=0, next_action=None, default_action=None, **device_kw):
        self.enable_action_pw = enable_action_pw
        self.alpha_observation, self.k_observation = alpha_observation, k_observation
        self.alpha_action, self.k_action = alpha_action, k_action
        self.next_action = next_action
        self._pomcp_kw = dict(eps=eps, max_depth=max_depth, c=c, tree_queries=tree_queries, rng=rng,
                              node_belief_updater=node_belief_updater, estimate_value=estimate_value, init_V=init_V,
                              init_N=init_N, default_action=default_action, **device_kw)

    def as_pomcp(self, pomdp):
        kw = dict(self._pomcp_kw)
        kw["dpw_solver"] = True  # src/solver.jl:161-274 is a different simulate() even where nothing widens
        if self.enable_action_pw:
            if self.next_action is not None and not isinstance(self.next_action, RandomActionGenerator):
                raise NotImplementedError("next_action: only RandomActionGenerator runs on device (src/actions.jl:38-41)")
            kw["dpw_action"] = (self.k_action, self.alpha_action)
        if pomdp.continuous_obs or pomdp.n_observations > self.k_observation:
            kw["dpw_observation"] = (self.k_observation, self.alpha_observation)
        # a discrete observation space of at most k_observation elements never widens (src/solver.jl:223)
        return POMCPSolver(**kw)


class RandomActionGenerator:
    """MCTS.RandomActionGenerator (next_action for the DPW solver, src/actions.jl:38-40)."""

    def __init__(self, rng=0):
        self.rng = rng


class PORollout(RolloutEstimator):
    """PORollout(solver, updater) (src/rollout.jl:12-16,48-51): partially observable rollout with an explicit
    rollout updater.  On device the updater is implied by the policy (VoidUpdater for RandomSolver, the
    previous-observation updater for FeedWhenCrying)."""

    def __init__(self, solver=None, updater=None):
        super().__init__(solver)
        self.updater = updater


PORolloutEstimator = PORollout  # deprecated alias, src/rollout.jl:17






def init_V(initializer, problem=None, h=None, action=None):
    """init_V(n::Number, ...) = Float64(n); callbacks are host functions (src/tree.jl:33-35)."""
    return float(initializer(problem, h, action)) if callable(initializer) else float(initializer)


def init_N(initializer, problem=None, h=None, action=None):
    """init_N(n::Number, ...) = Int(n) (src/tree.jl:42-44)."""
    return int(initializer(problem, h, action)) if callable(initializer) else int(initializer)


def sparse_actions(pomcp, pomdp, h=None, num_actions=0):
    """sparse_actions (src/actions.jl:10-33): the whole action space when num_actions <= 0 (the only branch
    the device path implements)."""
    if num_actions > 0:
        raise NotImplementedError("random action sub-sampling (src/actions.jl:15-29) is not supported on device")
    return list(range(pomdp.n_actions))






def convert_estimator(ev, solver, pomdp):
    """convert_estimator (src/rollout.jl:41-61): on device the estimator is resolved to a kind code."""
    s = POMCPSolver(estimate_value=ev)
    return s.c_params(abi.BELIEF_UNWEIGHTED).estimator


def extract_belief(rollout_updater, node):
    """extract_belief (src/rollout.jl:101-108): nothing for a VoidUpdater, the node's label for the
    previous-observation updater."""
    return None if rollout_updater is None else node.label


def estimate_value(estimator, pomdp, start_state, h=None, steps=0):
    """estimate_value(n::Number, ...) = Float64(n) (src/rollout.jl:10); rollout estimators run inside the search
    kernels (pomcp_rollout_batch evaluates them for given start states)."""
    if isinstance(estimator, (int, float)):
        return float(estimator)
    raise NotImplementedError("rollout estimators are evaluated on the device (POMCPPlanner.rollout_batch)")

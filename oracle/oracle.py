"""TEST INFRASTRUCTURE ONLY — ctypes binding of oracle/libpomcp_oracle.so.

The CPU restatement of the POMCP.jl hot path (oracle/pomcp_oracle.c) used as
the parity checker by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Product code (pomcp.jl_b200/) must never
import this module.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import pomcp_jl_b200._abi as abi
from pomcp_jl_b200.problems import obs_bits

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libpomcp_oracle.so")
_lib = None


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("pomcp_oracle.c", "pomcp_oracle.h", "Makefile")]
    src.append(os.path.join(_HERE, "..", "include", "pomcp_b200.h"))
    if not force and os.path.exists(_SO) and all(os.path.getmtime(_SO) >= os.path.getmtime(s) for s in src):
        return _SO
    subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i32, i64, u64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_uint64, C.c_double
        P = C.POINTER
        L.orc_create.restype = vp
        L.orc_create.argtypes = [P(abi.Problem), P(abi.SolverParams)]
        L.orc_destroy.argtypes = [vp]
        L.orc_destroy.restype = None
        L.orc_set_belief_particles.argtypes = [vp, vp, i64]
        L.orc_set_belief_initial.argtypes = [vp]
        L.orc_set_state_values.argtypes = [vp, vp, i64]
        L.orc_get_belief_particles.argtypes = [vp, vp, i64, P(i64)]
        L.orc_set_belief_exact2.argtypes = [vp, dbl]
        L.orc_get_belief_exact2.argtypes = [vp, P(dbl)]
        L.orc_update_exact2.argtypes = [vp, i32, u64]
        L.orc_bayes2.argtypes = [P(abi.Problem), dbl, i32, u64]
        L.orc_bayes2.restype = dbl
        L.orc_search.argtypes = [vp, i64, P(i32), vp, vp, i32]
        L.orc_search_replay.argtypes = [vp, i64, vp, i64, vp, i64, P(i32), vp, vp, i32, P(i64), P(i64)]
        L.orc_update_unweighted.argtypes = [vp, i32, u64]
        L.orc_pf_update_sir.argtypes = [vp, i32, u64, u64, i64, i32]
        L.orc_update_reinvigorated.argtypes = [vp, i32, u64, u64, i64, P(i32)]
        L.orc_resample_systematic_f64.argtypes = [vp, i64, dbl, i64, vp]
        L.orc_resample_systematic_fixed.argtypes = [vp, i64, dbl, i64, vp]
        L.orc_rollout_batch.argtypes = [P(abi.Problem), P(abi.SolverParams), vp, i64, i32, u64, vp, vp]
        L.orc_pf_weights.argtypes = [P(abi.Problem), vp, i64, i32, u64, u64, vp, vp]
        L.orc_get_root_stats.argtypes = [vp, P(i64), vp, vp, i32]
        L.orc_get_node_stats.argtypes = [vp, vp, vp, i32, P(i64), P(i64), vp, vp, i32]
        L.orc_dump_tree.argtypes = [vp, i64, P(i64), vp, vp, vp, vp]
        L.orc_get_counters.argtypes = [vp, P(abi.Counters)]
        L.orc_search_rootparallel.argtypes = [P(abi.Problem), P(abi.SolverParams), vp, i64, i32, i32, i64, P(i32),
                                              vp, vp, i32, P(abi.Counters)]
        L.orc_max_threads.restype = i32
        L.orc_philox4x32_10.argtypes = [vp, vp, vp]
        L.orc_philox4x32_10.restype = None
        for f in ("orc_det_log", "orc_det_exp", "orc_det_cos2pi"):
            getattr(L, f).argtypes = [dbl]
            getattr(L, f).restype = dbl
        L.orc_normal_from_uniforms.argtypes = [dbl, dbl]
        L.orc_normal_from_uniforms.restype = dbl
        L.orc_model_step.argtypes = [P(abi.Problem), vp, i32, vp, vp, vp, P(u64), P(dbl)]
        L.orc_obs_weight.argtypes = [P(abi.Problem), i32, vp, u64]
        L.orc_obs_weight.restype = dbl
        L.orc_is_terminal.argtypes = [P(abi.Problem), vp]
        _lib = L
    return _lib


def _vp(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def default_params(**kw):
    sp = abi.SolverParams()
    sp.eps, sp.max_depth, sp.c, sp.tree_queries = 0.01, 2 ** 63 - 1, 1.0, 100
    sp.seed, sp.init_V, sp.init_N = 0, 0.0, 0
    for k, v in kw.items():
        if not hasattr(sp, k):
            raise TypeError(k)
        setattr(sp, k, v)
    return sp


class OraclePlanner:
    """POMCPPlanner stand-in on the CPU oracle."""

    def __init__(self, pomdp, params):
        self.pomdp = pomdp
        self.nA = pomdp.n_actions
        self._p = pomdp.c_struct()
        self._sp = params
        self._h = lib().orc_create(C.byref(self._p), C.byref(self._sp))
        if not self._h:
            raise ValueError("orc_create failed (invalid problem/solver parameters)")

    def close(self):
        if self._h:
            lib().orc_destroy(self._h)
            self._h = None

    __del__ = close

    def set_belief_particles(self, states):
        states = np.ascontiguousarray(states, dtype=self.pomdp.state_dtype)
        return lib().orc_set_belief_particles(self._h, _vp(states), len(states))

    def set_belief_initial(self):
        return lib().orc_set_belief_initial(self._h)

    def set_belief_exact2(self, p1):
        return lib().orc_set_belief_exact2(self._h, float(p1))

    def belief_exact2(self):
        p = C.c_double(0.0)
        assert lib().orc_get_belief_exact2(self._h, C.byref(p)) == 0
        return p.value

    def update_exact2(self, a, o):
        return lib().orc_update_exact2(self._h, int(a), obs_bits(o))

    def set_state_values(self, values):
        values = np.ascontiguousarray(values, dtype=np.float64)
        return lib().orc_set_state_values(self._h, _vp(values), len(values))

    def belief_particles(self):
        n = C.c_int64(0)
        lib().orc_get_belief_particles(self._h, None, 0, C.byref(n))
        out = np.zeros(n.value, dtype=self.pomdp.state_dtype)
        lib().orc_get_belief_particles(self._h, _vp(out), n.value, C.byref(n))
        return out

    def search(self, tree_queries=-1):
        a = C.c_int32(-1)
        N = np.zeros(self.nA, dtype=np.int64)
        V = np.zeros(self.nA, dtype=np.float64)
        rc = lib().orc_search(self._h, tree_queries, C.byref(a), _vp(N), _vp(V), self.nA)
        return rc, a.value, N, V

    def search_replay(self, tree_queries, uniforms, returns):
        a = C.c_int32(-1)
        N = np.zeros(self.nA, dtype=np.int64)
        V = np.zeros(self.nA, dtype=np.float64)
        uu, ru = C.c_int64(0), C.c_int64(0)
        uniforms = np.ascontiguousarray(uniforms, dtype=np.float64)
        returns = np.ascontiguousarray(returns, dtype=np.float64)
        rc = lib().orc_search_replay(self._h, tree_queries, _vp(uniforms), len(uniforms), _vp(returns), len(returns),
                                     C.byref(a), _vp(N), _vp(V), self.nA, C.byref(uu), C.byref(ru))
        return rc, a.value, N, V, uu.value, ru.value

    def update_unweighted(self, a, o):
        return lib().orc_update_unweighted(self._h, a, obs_bits(o))

    def update_reinvigorated(self, a, o, seed, n_min):
        out = C.c_int32(-1)
        rc = lib().orc_update_reinvigorated(self._h, a, obs_bits(o), seed, n_min, C.byref(out))
        return rc, out.value

    def pf_update_sir(self, a, o, seed, n_out, mode=0):
        return lib().orc_pf_update_sir(self._h, a, obs_bits(o), seed, n_out, mode)

    def root_stats(self):
        rn = C.c_int64(0)
        N = np.zeros(self.nA, dtype=np.int64)
        V = np.zeros(self.nA, dtype=np.float64)
        lib().orc_get_root_stats(self._h, C.byref(rn), _vp(N), _vp(V), self.nA)
        return rn.value, N, V

    def node_stats(self, history):
        acts = np.array([h[0] for h in history], dtype=np.int32)
        obs = np.array([obs_bits(h[1]) for h in history], dtype=np.uint64)
        nN, npart = C.c_int64(0), C.c_int64(0)
        N = np.zeros(self.nA, dtype=np.int64)
        V = np.zeros(self.nA, dtype=np.float64)
        rc = lib().orc_get_node_stats(self._h, _vp(acts), _vp(obs), len(history), C.byref(nN), C.byref(npart),
                                      _vp(N), _vp(V), self.nA)
        assert rc == 0
        return nN.value, npart.value, N, V

    def dump_tree(self):
        n = C.c_int64(0)
        lib().orc_dump_tree(self._h, 0, C.byref(n), None, None, None, None)
        nn = n.value
        nO = getattr(self.pomdp, "n_observations", 0)
        node_N = np.zeros(nn, dtype=np.int64)
        act_N = np.zeros((nn, self.nA), dtype=np.int64)
        act_V = np.zeros((nn, self.nA), dtype=np.float64)
        child = np.full((nn, self.nA, max(nO, 1)), -1, dtype=np.int32) if nO else None
        lib().orc_dump_tree(self._h, nn, C.byref(n), _vp(node_N), _vp(act_N), _vp(act_V), _vp(child))
        return node_N, act_N, act_V, child

    def counters(self):
        c = abi.Counters()
        lib().orc_get_counters(self._h, C.byref(c))
        return c.as_dict()


def resample_systematic(w, r, n_out, fixed=True):
    w = np.ascontiguousarray(w, dtype=np.float64)
    idx = np.zeros(n_out, dtype=np.int64)
    f = lib().orc_resample_systematic_fixed if fixed else lib().orc_resample_systematic_f64
    rc = f(_vp(w), len(w), float(r), n_out, _vp(idx))
    return rc, idx


def rollout_batch(pomdp, states, steps, seed):
    states = np.ascontiguousarray(states, dtype=pomdp.state_dtype)
    p, sp = pomdp.c_struct(), default_params()
    ret = np.zeros(len(states), dtype=np.float64)
    ns = np.zeros(len(states), dtype=np.int32)
    rc = lib().orc_rollout_batch(C.byref(p), C.byref(sp), _vp(states), len(states), steps, seed, _vp(ret), _vp(ns))
    assert rc == 0, rc
    return ret, ns


def pf_weights(pomdp, states, a, o, seed):
    states = np.ascontiguousarray(states, dtype=pomdp.state_dtype)
    p = pomdp.c_struct()
    w = np.zeros(len(states), dtype=np.float64)
    nxt = np.zeros(len(states), dtype=pomdp.state_dtype)
    rc = lib().orc_pf_weights(C.byref(p), _vp(states), len(states), a, obs_bits(o), seed, _vp(w), _vp(nxt))
    assert rc == 0, rc
    return w, nxt


def model_step(pomdp, state, a, ut=(0.0, 0.0), uo=(0.0, 0.0)):
    p = pomdp.c_struct()
    s = np.array([state], dtype=pomdp.state_dtype)
    sp = np.zeros(1, dtype=pomdp.state_dtype)
    ut = np.array(ut, dtype=np.float64)
    uo = np.array(uo, dtype=np.float64)
    o, r = C.c_uint64(0), C.c_double(0.0)
    rc = lib().orc_model_step(C.byref(p), _vp(s), a, _vp(ut), _vp(uo), _vp(sp), C.byref(o), C.byref(r))
    assert rc == 0, rc
    return sp[0], o.value, r.value


def obs_weight(pomdp, a, next_state, o):
    p = pomdp.c_struct()
    s = np.array([next_state], dtype=pomdp.state_dtype)
    return lib().orc_obs_weight(C.byref(p), a, _vp(s), obs_bits(o))


def bayes2(pomdp, b, a, o):
    p = pomdp.c_struct()
    return lib().orc_bayes2(C.byref(p), float(b), int(a), obs_bits(o))


def philox(ctr, key):
    c = np.array(ctr, dtype=np.uint32)
    k = np.array(key, dtype=np.uint32)
    out = np.zeros(4, dtype=np.uint32)
    lib().orc_philox4x32_10(_vp(c), _vp(k), _vp(out))
    return tuple(int(x) for x in out)


def search_rootparallel(pomdp, params, belief_states, n_trees, n_threads, tree_queries):
    p = pomdp.c_struct()
    nA = pomdp.n_actions
    a = C.c_int32(-1)
    N = np.zeros(nA, dtype=np.int64)
    V = np.zeros(nA, dtype=np.float64)
    cs = abi.Counters()
    if belief_states is not None:
        belief_states = np.ascontiguousarray(belief_states, dtype=pomdp.state_dtype)
    rc = lib().orc_search_rootparallel(C.byref(p), C.byref(params), _vp(belief_states),
                                       0 if belief_states is None else len(belief_states), n_trees, n_threads,
                                       tree_queries, C.byref(a), _vp(N), _vp(V), nA, C.byref(cs))
    return rc, a.value, N, V, cs.as_dict()

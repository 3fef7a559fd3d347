"""Shared problem instances / solver settings for the parity tests."""
import numpy as np

import pomcp_jl_b200 as pj
from pomcp_jl_b200 import _abi as abi


def problems():
    return {
        "tiger": (pj.TigerPOMDP(), 1.0),
        "baby": (pj.BabyPOMDP(-5.0, -10.0), 10.0),
        "rs78": (pj.RockSamplePOMDP(7, 8, seed=7008), 20.0),
        "rs44": (pj.RockSamplePOMDP(4, 4, seed=44), 20.0),
        "lightdark": (pj.LightDark1D(), 1.0),
    }


def make_pair(O, pomdp, c, seed, mode="exact", belief_mode=abi.BELIEF_UNWEIGHTED, tree_queries=1000, **kw):
    """(gpu planner, oracle planner) with identical parameters."""
    solver = pj.POMCPSolver(c=c, rng=seed, tree_queries=tree_queries, search_mode=mode, **kw)
    gp = pj.solve(solver, pomdp, belief_mode=belief_mode)
    op = O.OraclePlanner(pomdp, solver.c_params(belief_mode))
    return gp, op


def random_states(pomdp, n, rng):
    if isinstance(pomdp, pj.LightDark1D):
        s = np.zeros(n, dtype=pomdp.state_dtype)
        s["y"] = rng.normal(2.0, 3.0, n)
        s["status"] = np.where(rng.random(n) < 0.1, -1, 0)
        return s
    if isinstance(pomdp, pj.RockSamplePOMDP):
        x = rng.integers(0, pomdp.n, n).astype(np.uint32)
        y = rng.integers(0, pomdp.n, n).astype(np.uint32)
        rocks = rng.integers(0, 1 << pomdp.k, n).astype(np.uint32)
        term = (rng.random(n) < 0.05).astype(np.uint32) * np.uint32(abi.RS_TERMINAL)
        return (x | (y << 4) | (rocks << 8) | term).astype(np.uint32)
    return rng.integers(0, 2, n).astype(np.uint32)

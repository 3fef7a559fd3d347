"""Problem descriptors for the models compiled into libpomcp_b200.so.

These stand in for the POMDPs.jl problem objects the reference is handed
(`solve(solver, pomdp)`, src/solver.jl:30-32): TigerPOMDP / BabyPOMDP /
LightDark1D as in POMDPModels (test/runtests.jl:4, notebooks/Minimal_Example.ipynb:68-146)
and RockSample(n,k) (BASELINE.json configs; standard Smith & Simmons definition).
Each object only carries the POD parameter block of include/pomcp_b200.h.
"""
import ctypes as C
import struct

from . import _abi


def philox4x32_10(ctr, key):
    """Philox4x32-10 (host-side twin of csrc/philox.cuh), python ints."""
    c0, c1, c2, c3 = ctr
    k0, k1 = key
    for _ in range(10):
        p0 = 0xD2511F53 * c0
        p1 = 0xCD9E8D57 * c2
        c0, c1, c2, c3 = ((p1 >> 32) ^ c1 ^ k0) & 0xFFFFFFFF, p1 & 0xFFFFFFFF, \
            ((p0 >> 32) ^ c3 ^ k1) & 0xFFFFFFFF, p0 & 0xFFFFFFFF
        k0 = (k0 + 0x9E3779B9) & 0xFFFFFFFF
        k1 = (k1 + 0xBB67AE85) & 0xFFFFFFFF
    return c0, c1, c2, c3


def obs_bits(o):
    """Observation -> the uint64 the C ABI takes (index, or IEEE bits of a float)."""
    if isinstance(o, float):
        return struct.unpack("<Q", struct.pack("<d", o))[0]
    return int(o)


def obs_from_bits(problem, bits):
    if problem.continuous_obs:
        return struct.unpack("<d", struct.pack("<Q", int(bits)))[0]
    return int(bits)


class POMDP:
    problem_id = None
    continuous_obs = False

    def c_struct(self) -> _abi.Problem:
        raise NotImplementedError

    @property
    def n_actions(self):
        raise NotImplementedError

    @property
    def state_dtype(self):
        import numpy as np
        return np.dtype(np.uint32)

    @property
    def discount(self):
        return self._discount


class TigerPOMDP(POMDP):
    """POMDPModels.TigerPOMDP: S=Bool(tiger left), A={0 listen,1 open-left,2 open-right}, O=Bool."""
    problem_id = _abi.TIGER

    def __init__(self, r_listen=-1.0, r_findtiger=-100.0, r_escapetiger=10.0, p_listen_correctly=0.85,
                 discount_factor=0.95):
        self.r_listen, self.r_findtiger, self.r_escapetiger = r_listen, r_findtiger, r_escapetiger
        self.p_listen_correctly = p_listen_correctly
        self._discount = discount_factor

    n_actions = 3
    n_observations = 2
    initial_p1 = 0.5  # initial_state_distribution: the tiger is behind either door

    def c_struct(self):
        p = _abi.Problem()
        p.problem_id = self.problem_id
        p.discount = self._discount
        p.tiger_r_listen, p.tiger_r_findtiger = self.r_listen, self.r_findtiger
        p.tiger_r_escapetiger, p.tiger_p_listen_correctly = self.r_escapetiger, self.p_listen_correctly
        return p


class BabyPOMDP(POMDP):
    """POMDPModels.BabyPOMDP(r_feed, r_hungry): S=A=O=Bool (notebooks/Basic_Usage.ipynb:37)."""
    problem_id = _abi.BABY

    def __init__(self, r_feed=-5.0, r_hungry=-10.0, p_become_hungry=0.1, p_cry_when_hungry=0.8,
                 p_cry_when_not_hungry=0.1, discount_factor=0.9):
        self.r_feed, self.r_hungry = float(r_feed), float(r_hungry)
        self.p_become_hungry, self.p_cry_when_hungry = p_become_hungry, p_cry_when_hungry
        self.p_cry_when_not_hungry = p_cry_when_not_hungry
        self._discount = discount_factor

    n_actions = 2
    n_observations = 2
    initial_p1 = 0.0  # initial_state_distribution = BoolDistribution(0.0): not hungry

    def c_struct(self):
        p = _abi.Problem()
        p.problem_id = self.problem_id
        p.discount = self._discount
        p.baby_r_feed, p.baby_r_hungry = self.r_feed, self.r_hungry
        p.baby_p_become_hungry = self.p_become_hungry
        p.baby_p_cry_when_hungry = self.p_cry_when_hungry
        p.baby_p_cry_when_not_hungry = self.p_cry_when_not_hungry
        return p


class RockSamplePOMDP(POMDP):
    """RockSample(n,k).  Rocks default to k distinct cells drawn with Philox(seed) without
    replacement (seed 7008 for (7,8), 11011 for (11,11)); start = (0, n//2)."""
    problem_id = _abi.ROCKSAMPLE

    def __init__(self, n=7, k=8, rocks=None, seed=None, start=None, half_eff_dist=20.0, r_good=10.0, r_bad=-10.0,
                 r_exit=10.0, discount_factor=0.95):
        if not (1 <= n <= 15 and 0 <= k <= _abi.MAX_ROCKS and k <= n * n):
            raise ValueError("RockSample: need 1<=n<=15, 0<=k<=16")
        self.n, self.k = n, k
        if rocks is None:
            if seed is None:
                seed = n * 1000 + k
            cells = list(range(n * n))
            rocks = []
            for i in range(k):
                w = philox4x32_10((i, 0, 0, 0), (seed & 0xFFFFFFFF, (seed >> 32) & 0xFFFFFFFF))[0]
                pos = cells.pop((w * len(cells)) >> 32)
                rocks.append((pos % n, pos // n))
        self.rocks = [tuple(r) for r in rocks]
        assert len(self.rocks) == k and len(set(self.rocks)) == k
        self.start = tuple(start) if start is not None else (0, n // 2)
        self.half_eff_dist, self.r_good, self.r_bad, self.r_exit = half_eff_dist, r_good, r_bad, r_exit
        self._discount = discount_factor

    @property
    def n_actions(self):
        return 5 + self.k

    n_observations = 3

    def pack_state(self, x, y, rocks_bits, terminal=False):
        return (x & 15) | ((y & 15) << 4) | ((rocks_bits & 0xFFFF) << 8) | (_abi.RS_TERMINAL if terminal else 0)

    def c_struct(self):
        p = _abi.Problem()
        p.problem_id = self.problem_id
        p.discount = self._discount
        p.rs_n, p.rs_k = self.n, self.k
        for i, (x, y) in enumerate(self.rocks):
            p.rs_rock_x[i], p.rs_rock_y[i] = x, y
        p.rs_start_x, p.rs_start_y = self.start
        p.rs_half_eff_dist = self.half_eff_dist
        p.rs_r_good, p.rs_r_bad, p.rs_r_exit = self.r_good, self.r_bad, self.r_exit
        return p


class LightDark1D(POMDP):
    """LightDark1D exactly as defined inside the reference (notebooks/Minimal_Example.ipynb:68-146):
    actions index {0,1,2} = moves {-1, 0(stop), +1}; observation is a Float64."""
    problem_id = _abi.LIGHTDARK1D
    continuous_obs = True

    def __init__(self, discount_factor=0.9, correct_r=10.0, incorrect_r=-10.0, init_mean=2.0, init_std=3.0):
        self._discount = discount_factor
        self.correct_r, self.incorrect_r = correct_r, incorrect_r
        self.init_mean, self.init_std = init_mean, init_std

    n_actions = 3
    n_observations = 0

    @property
    def state_dtype(self):
        import numpy as np
        return np.dtype([("status", np.int64), ("y", np.float64)])

    def c_struct(self):
        p = _abi.Problem()
        p.problem_id = self.problem_id
        p.discount = self._discount
        p.ld_correct_r, p.ld_incorrect_r = self.correct_r, self.incorrect_r
        p.ld_init_mean, p.ld_init_std = self.init_mean, self.init_std
        return p


def state_ptr(arr):
    return arr.ctypes.data_as(C.c_void_p)

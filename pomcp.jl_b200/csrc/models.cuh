// Generative models compiled into the library, one specialisation per concrete POMDP.
// They replace the POMDPs.jl callbacks the reference's hot loop calls down into:
// generate_sor (src/solver.jl:126), isterminal (src/solver.jl:49,83), discount, actions
// (src/actions.jl:30-31), and — for the SIR filter — generate_s / observation / pdf.
// Transition draws come first, observation draws second (POMDPs.jl generate_sor default).
#pragma once
#include <cstdint>
#include "../../include/pomcp_b200.h"
#include "detmath.cuh"
#include "philox.cuh"

namespace pomcp {

struct LDState {
  long long status;  // < 0 terminal (notebooks/Minimal_Example.ipynb:77)
  double y;
};

// Device-side parameter block, passed by value (__grid_constant__) to every kernel.
struct DevModel {
  int pid, nA, nO, state_bytes;
  double gamma;
  double t_r_listen, t_r_find, t_r_escape, t_p_correct;
  double b_r_feed, b_r_hungry, b_p_hungry, b_p_cry_h, b_p_cry_nh;
  unsigned long long b_thr_hungry;  // word < thr  <=>  word*2^-32 < b_p_hungry
  int rs_n, rs_k;
  unsigned rs_start;  // packed x | y<<4
  double rs_r_good, rs_r_bad, rs_r_exit;
  double rs_rtab[4];  // rollout reward by event: 0 none, 1 exit, 2 sampled a bad rock, 3 sampled a good rock
  const double* rs_eta;  // device table [cell(y*16+x)][rock]: P(check correct)
  signed char rock_at[256];
  double ld_correct, ld_incorrect, ld_mean, ld_std;
};

template <class U>
__device__ __forceinline__ bool draw_lt(U u, double p, unsigned long long thr);
template <>
__device__ __forceinline__ bool draw_lt<double>(double u, double p, unsigned long long) { return u < p; }
template <>
__device__ __forceinline__ bool draw_lt<uint32_t>(uint32_t w, double, unsigned long long thr) {
  return (unsigned long long)w < thr;
}

__device__ __forceinline__ double ld_sigma(double x) {  // notebooks/Minimal_Example.ipynb:89
  return fabs(x - 5.0) / 1.4142135623730951 + 1e-2;
}

template <int PID>
struct ModelT;

// ---------------------------------------------------------------- Tiger (POMDPModels.TigerPOMDP)
template <>
struct ModelT<POMCP_TIGER> {
  using State = uint32_t;  // 1: tiger behind the left door
  static constexpr int DPS = 2;  // Philox words per rollout step: action, transition
  static constexpr bool HASHED = false;
  static constexpr int NO = 2;
  static constexpr bool T_DRAW = true;  // the transition consumes a uniform
  static constexpr bool TWO_STATE = true;
  // P(s' = 1 | s, a): listening keeps the tiger where it is, opening a door resets it uniformly
  __device__ __forceinline__ static double trans_p1(const DevModel&, State s, int a) { return a == 0 ? (s ? 1.0 : 0.0) : 0.5; }
  __device__ __forceinline__ static bool terminal(const DevModel&, State) { return false; }
  template <bool OBS, class U>
  __device__ __forceinline__ static void step(const DevModel& m, State s, int a, U ut0, double uo0, double,
                                              State& sp, uint64_t& o, double& r) {
    sp = s;
    o = 0;
    if (a == 0) {
      r = m.t_r_listen;
      if (OBS) o = (uo0 < m.t_p_correct) ? (uint64_t)s : (uint64_t)(1u - s);
    } else {
      uint32_t opened_tiger = (a == 1) ? s : (1u - s);
      r = opened_tiger ? m.t_r_find : m.t_r_escape;
      sp = draw_lt<U>(ut0, 0.5, 0x80000000ull) ? 1u : 0u;
      if (OBS) o = (uo0 < 0.5) ? 1u : 0u;
    }
  }
  __device__ __forceinline__ static double obs_weight(const DevModel& m, int a, State sp, uint64_t obs) {
    if (a == 0) return ((uint64_t)sp == obs) ? m.t_p_correct : 1.0 - m.t_p_correct;
    return 0.5;
  }
  __device__ __forceinline__ static State initial(const DevModel&, double u0, double) { return u0 < 0.5 ? 1u : 0u; }
};

// ---------------------------------------------------------------- Baby (POMDPModels.BabyPOMDP)
template <>
struct ModelT<POMCP_BABY> {
  using State = uint32_t;  // 1: hungry
  static constexpr int DPS = 2;
  static constexpr bool HASHED = false;
  static constexpr int NO = 2;
  static constexpr bool T_DRAW = true;  // the transition consumes a uniform
  static constexpr bool TWO_STATE = true;
  // P(hungry' | s, a): feeding sates, a hungry baby stays hungry, otherwise it becomes hungry w.p. p_become_hungry
  __device__ __forceinline__ static double trans_p1(const DevModel& m, State s, int a) {
    return a ? 0.0 : (s ? 1.0 : m.b_p_hungry);
  }
  __device__ __forceinline__ static bool terminal(const DevModel&, State) { return false; }
  template <bool OBS, class U>
  __device__ __forceinline__ static void step(const DevModel& m, State s, int a, U ut0, double uo0, double,
                                              State& sp, uint64_t& o, double& r) {
    r = (s ? m.b_r_hungry : 0.0) + (a ? m.b_r_feed : 0.0);
    if (a) sp = 0u;
    else if (s) sp = 1u;
    else sp = draw_lt<U>(ut0, m.b_p_hungry, m.b_thr_hungry) ? 1u : 0u;
    o = 0;
    if (OBS) {
      double pc = sp ? m.b_p_cry_h : m.b_p_cry_nh;
      o = (uo0 < pc) ? 1u : 0u;
    }
  }
  __device__ __forceinline__ static double obs_weight(const DevModel& m, int, State sp, uint64_t obs) {
    double pc = sp ? m.b_p_cry_h : m.b_p_cry_nh;
    return obs ? pc : 1.0 - pc;
  }
  __device__ __forceinline__ static State initial(const DevModel&, double, double) { return 0u; }  // BoolDistribution(0.0)
};

// ---------------------------------------------------------------- RockSample(n,k)
template <>
struct ModelT<POMCP_ROCKSAMPLE> {
  using State = uint32_t;  // bits 0-3 x, 4-7 y, 8-23 rocks (1 good), bit 31 terminal
  static constexpr int DPS = 1;
  static constexpr bool HASHED = false;
  static constexpr int NO = 3;
  static constexpr bool T_DRAW = false;  // deterministic transitions
  static constexpr bool TWO_STATE = false;
  __device__ __forceinline__ static bool terminal(const DevModel&, State s) { return (s & POMCP_RS_TERMINAL) != 0; }
  template <bool OBS, class U>
  __device__ __forceinline__ static void step(const DevModel& m, State st, int a, U, double uo0, double, State& sp,
                                              uint64_t& o, double& r) {
    int x = (int)(st & 15u), y = (int)((st >> 4) & 15u);
    o = 2;  // none
    // moves without branches: lanes of a rollout warp hold different actions, so every arm of an
    // if-chain would be executed by the warp on every step
    const int nmax = m.rs_n - 1;
    y = min(max(y + (int)(a == 0) - (int)(a == 2), 0), nmax);
    const int nx = x + (int)(a == 1) - (int)(a == 3);
    const bool exits = nx > nmax;  // only a == 1 on the east edge
    x = min(max(nx, 0), nmax);
    st |= exits ? POMCP_RS_TERMINAL : 0u;
    r = exits ? m.rs_r_exit : 0.0;
    if (a == 4) {
      int ri = m.rock_at[y * 16 + x];
      if (ri >= 0) {
        uint32_t good = (st >> (8 + ri)) & 1u;
        r = good ? m.rs_r_good : m.rs_r_bad;
        st &= ~(good << (8 + ri));
      }
    } else if (a >= 5) {
      if (OBS) {
        int ri = a - 5;
        int good = (int)((st >> (8 + ri)) & 1u);
        double e = m.rs_eta[(y * 16 + x) * POMCP_MAX_ROCKS + ri];
        int correct = uo0 < e;
        o = (uint64_t)((good == correct) ? 0 : 1);
      } else {
        o = 0;
      }
    }
    sp = (st & ~0xffu) | (uint32_t)x | ((uint32_t)y << 4);
  }
  __device__ __forceinline__ static double obs_weight(const DevModel& m, int a, State sp, uint64_t obs) {
    if (a < 5) return obs == 2 ? 1.0 : 0.0;
    if (obs > 1) return 0.0;
    int x = (int)(sp & 15u), y = (int)((sp >> 4) & 15u), ri = a - 5;
    int good = (int)((sp >> (8 + ri)) & 1u);
    double e = m.rs_eta[(y * 16 + x) * POMCP_MAX_ROCKS + ri];
    int says_good = obs == 0;
    return (good == says_good) ? e : 1.0 - e;
  }
  __device__ __forceinline__ static State initial(const DevModel& m, double u0, double) {
    uint32_t rocks = (uint32_t)(u0 * (double)(1u << m.rs_k));
    return m.rs_start | (rocks << 8);
  }
};

// ---------------------------------------------------------------- LightDark1D (notebooks/Minimal_Example.ipynb:68-146)
template <>
struct ModelT<POMCP_LIGHTDARK1D> {
  using State = LDState;
  static constexpr int DPS = 1;
  static constexpr bool HASHED = true;  // Float64 observations: children live in a hash table
  static constexpr int NO = 0;
  static constexpr bool T_DRAW = false;  // deterministic transitions
  static constexpr bool TWO_STATE = false;
  __device__ __forceinline__ static bool terminal(const DevModel&, const State& s) { return s.status < 0; }
  template <bool OBS, class U>
  __device__ __forceinline__ static void step(const DevModel& m, const State& s, int a, U, double uo0, double uo1,
                                              State& sp, uint64_t& o, double& r) {
    sp = s;
    r = 0.0;
    o = 0;
    if (s.status < 0) {
      r = 0.0;
    } else if (a == 1) {
      r = (fabs(s.y) < 1.0) ? m.ld_correct : m.ld_incorrect;
      sp.status = -1;
    } else {
      sp.y = s.y + (double)(a - 1);
    }
    if (OBS) {
      double z = det_normal(uo0, uo1);
      o = d2u(sp.y + z * ld_sigma(sp.y));
    }
  }
  __device__ __forceinline__ static double obs_weight(const DevModel&, int, const State& sp, uint64_t obs) {
    double x = u2d(obs), sd = ld_sigma(sp.y);
    double d = x - sp.y;
    return det_exp(-(d * d) / (2.0 * (sd * sd))) / (sd * 2.5066282746310002);
  }
  __device__ __forceinline__ static State initial(const DevModel& m, double u0, double u1) {
    State s;
    s.status = 0;
    s.y = m.ld_mean + det_normal(u0, u1) * m.ld_std;
    return s;
  }
};

// Exact Bayes filter of a two-state problem (POMDPModels' BabyBeliefUpdater / the DiscreteUpdater of Tiger), the
// node_belief_updater of notebooks/Sanity_Checks.ipynb:47,70 (`updater(problem)`; used at src/solver.jl:138 and
// src/updater.jl:33): b = P(s = 1), b' proportional to O(o | a, s') * sum_s T(s' | s, a) b(s).  The operation order
// is part of the spec (the oracle evaluates the same expression; Baby reproduces the notebooks' belief chain).
template <int PID>
__device__ __forceinline__ double bayes2(const DevModel& m, double b, int a, uint64_t obs) {
  using M = ModelT<PID>;
  if constexpr (M::TWO_STATE) {
    const double p1 = b * M::trans_p1(m, 1u, a) + (1.0 - b) * M::trans_p1(m, 0u, a);
    const double n1 = M::obs_weight(m, a, 1u, obs) * p1;
    const double n0 = M::obs_weight(m, a, 0u, obs) * (1.0 - p1);
    const double den = n0 + n1;
    return den > 0.0 ? n1 / den : p1;
  } else {
    return b;
  }
}

}  // namespace pomcp

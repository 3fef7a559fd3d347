// K1/K2/K3/K4/K5 — the search core: root sampling, warp-per-path descent with UCB1 selection
// and child lookup/insert on the SoA node pool, one-thread-per-trajectory random rollouts reduced
// by warp shuffles, backup with atomics.  Replaces search()/simulate() (src/solver.jl:41-157),
// the tree node types (src/tree.jl:7-26) and rollout() (src/rollout.jl:77-84) of the reference.
//
// Node pool in HBM (struct of arrays, ids are u32, id 0 is the first root):
//   node_N[n]            visits of belief node n, counted on arrival (== reference N once drained)
//   node_flags[n]        bit0: children expanded (src/solver.jl:87-94)
//   act_N[n*A+a]         visits of action node (n,a), counted at descent (in-flight included)
//   act_M[n*A+a]         completed backups (batched mode)
//   act_V[n*A+a]         exact mode: running mean V; batched mode: W = sum of returns
//   child[(n*A+a)*nO+o]  discrete observations: child belief node id, 0 = none
//   ht[] / node_pslot[] / node_obs[]   continuous observations: open-addressed (slot,obs)->id
//   log_node[] / log_state[]           append-only particle log = push!(hao.B, sp) (src/solver.jl:146-148)
#pragma once
#include <type_traits>
#include <cstdint>
#include "models.cuh"

namespace pomcp {

// Direct child table (discrete observations): bit 31 of a child id says that the worker which
// created the node also expands it (it was not cut off there) — later visitors need not read the
// node's expansion flag, and the creator needs no claim round trip.
constexpr uint32_t CHILD_EXPANDED = 0x80000000u;
constexpr uint32_t CHILD_ID_MASK = 0x7fffffffu;

struct DevCtr {
  unsigned long long node_count, log_count;
  unsigned long long sims, skips, rollouts, rollout_steps, tree_steps, max_depth;
  unsigned long long scan_ticket;
  unsigned long long next_query, worker_ticket;
  unsigned int pool_overflow, log_overflow, replay_overflow, any_nonterminal;
  unsigned long long dbg[8];
};

struct Pool {
  uint32_t* node_N;
  uint32_t* node_flags;
  uint32_t* act_N;
  uint32_t* act_M;
  double* act_V;
  uint32_t* child;
  uint32_t* ht;
  uint32_t* node_pslot;
  unsigned long long* node_obs;
  uint32_t* log_node;
  void* log_state;
  // observation progressive widening (POMCPDPWSolver, continuous observations): children / generation events of
  // an action node, and the (slot, event) -> (child, state) table the widened step draws from
  double* node_bel;      // exact two-state node beliefs P(s = 1) (NULL unless belief_mode = EXACT2; src/solver.jl:138)
  uint32_t* node_amask;  // action widening: bit a = ActNode a exists (NULL unless dpw_action)
  uint32_t* act_nc;
  uint32_t* act_ng;
  unsigned long long* ev_key;
  uint32_t* ev_node;
  void* ev_state;
  unsigned long long ev_mask;
  DevCtr* ctr;
  unsigned long long ht_mask, log_cap;
  uint32_t max_nodes;
};

struct TreeCtl;

struct SearchCfg {
  double c, init_V, log_ratio, est_value;
  long long max_depth, init_N;
  const double* gpow;  // gamma^d, d = 0..levels
  const double* state_values;  // FOValue table (POMCP_EST_STATE_VALUE), n_state_values entries
  long long n_state_values;
  int levels;          // first depth at which gamma^depth < eps (src/solver.jl:82), capped by max_depth
  int belief_unweighted, estimator;
  int dpw;             // observation progressive widening on (src/solver.jl:223-262)
  double k_obs, alpha_obs;
  int dpa;             // action progressive widening on (src/solver.jl:172-186)
  double k_act, alpha_act;
  int exact2;          // node beliefs are exact two-state distributions (Pool::node_bel), roots are sampled from them
  int dpw_solver;      // simulate() of POMCPDPWSolver (src/solver.jl:161-274): the leaf estimate gets `depth` as its step
                       // budget (:182,:199) and an untried action scores V while total_N <= 1 (:210)
  int rollouts;        // R rollouts averaged per leaf, lane r runs rollout r
  int team_warps;      // batched mode: warps per worker; warp w runs rollouts 32 w .. 32 w + 31 of every leaf
  uint32_t key0, key1, rank, epoch;
  const void* bel;  // root belief particles (NULL: the problem's initial distribution)
  long long n_bel;
  uint32_t root;
  uint32_t root_N0;  // root visits when the search started
  TreeCtl* trees;    // batched mode: per-tree parameters (n_trees entries); key / rank / epoch / bel / root above are tree 0's
  int n_trees;
  // asynchronous workers: in flight = clamp(root visits / ramp_div, inflight_min, inflight_max)
  int inflight_min, inflight_max, ramp_div;
  int path_smem;  // batched mode: the worker's recorded path lives in shared memory (else in PathBuf)
  int path_bytes;      // bytes of one worker's path in dynamic shared memory (16-byte multiple)
  int eta_off, eta_n;  // RockSample: byte offset / entry count of the sensor thresholds in dynamic shared memory
  const uint32_t* eta_thr;  // [(y*n+x)*k+rock]: check is correct iff obs word <= threshold  (== u < eta)
  // replay (exact mode only)
  const double* uni;
  long long n_uni;
  const double* rets;
  long long n_rets;
};

// Several trees in one launch (root-parallel replicas of one belief, or independent episodes): everything that
// differs from tree to tree lives in one TreeCtl per tree in global memory; the rest of SearchCfg is shared.
// A worker belongs to one tree for its whole life: tree = worker id % n_trees, and the admission ramp, the
// query tickets and the in-flight cap are per tree.
struct TreeCtl {
  unsigned long long next_query;  // query ticket of this tree (zeroed by the host before every search)
  const void* bel;                // root belief particles (NULL: the problem's initial distribution)
  long long n_bel;
  uint32_t root, root_N0;         // root node id, root visits when the search started
  uint32_t key0, key1, rank, epoch;  // Philox key / stream id / epoch of this tree
  uint32_t any_nonterminal;       // a query of this search ran a simulation (src/solver.jl:56-58)
  uint32_t _pad;
  // particle log of this tree (unweighted mode): entries [log_base, log_base + log_cap) of Pool::log_*, cursor *log_ctr
  // (one tree: the whole log and DevCtr::log_count; episodes: one segment and one cursor per tree)
  unsigned long long* log_ctr;
  unsigned long long log_base, log_cap;
};
// A worker's copy of its tree's parameters, in shared memory (one per team; the helper warps read the keys).
struct TreeView {
  const void* bel;
  long long n_bel;
  TreeCtl* ctl;
  unsigned long long* log_ctr;
  unsigned long long log_base, log_cap;
  uint32_t root, root_N0, key0, key1, tagrank, epoch;
};
struct RngKey {
  uint32_t key0, key1, tagrank, epoch;  // tagrank = rank << 8
};

// Per-worker scratch: path[level][slot], slot = worker index
struct PathBuf {
  uint32_t* path_slot;
  uint32_t* path_node;
  double* path_r;
  void* path_state;
  int stride;
};

struct ReplayCursor {
  long long ucur, rcur;
};

struct LocalCtr {
  unsigned long long sims, skips, rollouts, rollout_steps, tree_steps, max_depth;
  // batched workers: SM cycles spent per phase (pomcp_worker_cycles): descent, insert, rollout, backup
  unsigned long long cyc[8];  // [4..7]: developer probes (-DPOMCP_PROBE)
};

struct SearchResult {
  int status, action;
  long long root_N;
  long long N[POMCP_MAX_ACTIONS];
  double V[POMCP_MAX_ACTIONS];
  double merge[3 * POMCP_MAX_ACTIONS];  // [N_a..., N_a*V_a..., exists_a...] for the root-parallel all-reduce
  long long ucur, rcur;
};

__device__ __forceinline__ unsigned long long ld_volatile_u64g(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}
__device__ __forceinline__ uint64_t mix64(uint64_t x) {
  x ^= x >> 30; x *= 0xbf58476d1ce4e5b9ull; x ^= x >> 27; x *= 0x94d049bb133111ebull; x ^= x >> 31;
  return x;
}

template <bool REPLAY>
__device__ __forceinline__ void get_draws(const SearchCfg& cfg, ReplayCursor& rc, uint32_t tag, uint32_t idx,
                                          uint32_t q, double out[4]) {
  if (REPLAY) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (rc.ucur < cfg.n_uni) out[i] = cfg.uni[rc.ucur]; else out[i] = 0.0;
      rc.ucur++;
    }
  } else {
    Philox4 p = philox4x32_10(idx, tag | (cfg.rank << 8), q, cfg.epoch, cfg.key0, cfg.key1);
#pragma unroll
    for (int i = 0; i < 4; ++i) out[i] = u01(p.w[i]);
  }
}

// CHECK_DISC: keep RolloutSimulator's `disc > eps` test (eps = 0).  Inside the tree
// steps <= ceil(log(eps)/log(gamma)), so disc >= eps * gamma > 0 and the test can never fire.
// Both rollout loops below are branch-free inside a Philox block: every step of the block is
// computed by every lane, and a lane that has finished (terminal state or step budget) only adds
// +0.0 to its return.  That leaves one basic block per Philox block, in which the next block's
// Philox rounds (independent of the trajectory) are interleaved with the dependent chain of the
// steps — a single trajectory is a latency chain, the issue slots next to it are free.
template <int PID, bool CHECK_DISC>
__device__ __forceinline__ double rollout_generic(const DevModel& m, typename ModelT<PID>::State s, long long steps,
                                                  uint32_t k0, uint32_t k1, uint32_t tagword, uint32_t q,
                                                  uint32_t epoch, int& nsteps) {
  using M = ModelT<PID>;
  constexpr int DPS = M::DPS;
  const double gamma = m.gamma;
  double disc = 1.0, total = 0.0;
  const int nstep_max = (int)(steps < 0x7fffffffll ? steps : 0x7fffffffll);
  int step = 0;
  bool live = !M::terminal(m, s) && nstep_max > 0;
  Philox4 p = philox4x32_10(0u, tagword, q, epoch, k0, k1);
  for (uint32_t blk = 0; live; ++blk) {
    const Philox4 pn = philox4x32_10(blk + 1u, tagword, q, epoch, k0, k1);
#pragma unroll
    for (int j = 0; j < 4 / DPS; ++j) {
      uint32_t wa = p.w[j * DPS];
      uint32_t wt = p.w[j * DPS + (DPS - 1)];
      int a = (int)__umulhi(wa, (uint32_t)m.nA);  // rand(rng, actions): floor(u * nA)
      typename M::State sp;
      uint64_t o;
      double r;
      M::template step<false, uint32_t>(m, s, a, wt, 0.0, 0.0, sp, o, r);
      total += disc * (live ? r : 0.0);
      s = sp;
      disc *= gamma;
      step += live ? 1 : 0;
      live = live && (!CHECK_DISC || disc > 0.0) && !M::terminal(m, s) && step < nstep_max;
    }
    p = pn;
  }
  nsteps = step;
  return total;
}

// RockSample lookup tables staged in shared memory: 32 lanes index them with 32 different cells,
// which the constant cache would serialise (one address per access) and shared memory does not.
struct RsTab {
  double rtab[4];           // reward by event: 0 none, 1 exit, 2 sampled a bad rock, 3 sampled a good rock
  double rtab2[6];          // the same by the rollout table's event field (byte offset 0 none, 16 exit, 32 bad, 40 good)
  signed char rock_at[256]; // rock index at cell y*16+x, -1 none
  // Rollout transition table: row = cell (y*16+x; row 255 = off the grid / finished, absorbing), 64 bytes per row,
  // column = action (16 columns; actions >= 15 share column 15: from 5 on every action is a sensing action, which
  // neither moves nor pays).  Entry, laid out so that a rollout step needs no field extraction at all:
  //   bits 0-1   zero                 bits 6-13  next row, i.e. bits 6-13 of the next entry's byte offset
  //   bits 2-3   zero                 bits 14-29 mask of the rock the action samples (bit 14 + i), 0 = none
  //   bits 4-5   event: byte offset / 16 into rtab2 (0 none, 1 exit, 2 sampled)
  //   bit 31     the row this entry belongs to is on the grid (counts live steps)
  // The next lookup address is (entry & RS_ROW_MASK) | (4 * action): one LOP3, since 4 * action < 64 fits bits 2-5.
  uint32_t mv[256 * 16];
};
constexpr uint32_t RS_ROW_MASK = 0x3fc3u;      // row field + the two always-zero low bits
constexpr uint32_t RS_DEAD_ROW = 255u * 64u;
constexpr int RS_ROCK_SHIFT = 14;
__device__ __forceinline__ void rs_tab_init(RsTab* t, const DevModel& m) {
  const int nmax = m.rs_n - 1;
  for (int i = threadIdx.x; i < 256; i += blockDim.x) t->rock_at[i] = m.rock_at[i];
  for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) {
    const int c = i >> 4, a = i & 15, x = c & 15, y = c >> 4;
    uint32_t next = (uint32_t)c, ev = 0, mask = 0, alive = 1;
    if (c == 255 || x > nmax || y > nmax) {
      next = 255;  // absorbing
      alive = 0;
    } else if (a == 0) {
      next = (uint32_t)(min(y + 1, nmax) * 16 + x);
    } else if (a == 1) {  // east: off the grid at the right edge
      if (x + 1 > nmax) { next = 255; ev = 1; } else next = (uint32_t)(y * 16 + x + 1);
    } else if (a == 2) {
      next = (uint32_t)(max(y - 1, 0) * 16 + x);
    } else if (a == 3) {
      next = (uint32_t)(y * 16 + max(x - 1, 0));
    } else if (a == 4 && m.rock_at[c] >= 0) {
      ev = 2;
      mask = 1u << m.rock_at[c];
    }
    t->mv[i] = (next << 6) | (ev << 4) | (mask << RS_ROCK_SHIFT) | (alive << 31);
  }
  if (threadIdx.x < 4) t->rtab[threadIdx.x] = m.rs_rtab[threadIdx.x];
  if (threadIdx.x < 6) {
    const int k = threadIdx.x;
    t->rtab2[k] = k == 2 ? m.rs_rtab[1] : k == 4 ? m.rs_rtab[2] : k == 5 ? m.rs_rtab[3] : 0.0;
  }
  __syncthreads();
}

__device__ __forceinline__ uint32_t shr_clamped(uint32_t x, uint32_t n) {  // 0 for n >= 32
  uint32_t y;
  asm("shr.u32 %0, %1, %2;" : "=r"(y) : "r"(x), "r"(n));
  return y;
}
__device__ __forceinline__ uint32_t shl_clamped(uint32_t x, uint32_t n) {  // 0 for n >= 32
  uint32_t y;
  asm("shl.b32 %0, %1, %2;" : "=r"(y) : "r"(x), "r"(n));
  return y;
}
__device__ __forceinline__ int sbfe2(uint32_t x, int pos) {  // signed 2-bit field
  int y;
  asm("bfe.s32 %0, %1, %2, 2;" : "=r"(y) : "r"(x), "r"(pos));
  return y;
}

// RockSample rollouts: the same draws, transitions, rewards and fp64 operation order as
// ModelT<ROCKSAMPLE>::step<false> (bit-identical returns, checked against the oracle through
// pomcp_rollout_batch), arranged for the issue slots (every pipe of a scheduler takes one warp
// instruction per two cycles, and the search kernel is bound by exactly that).  A step is: 16 bits of a Philox
// word -> 4 * action (one IMAD.HI: floor(4 u nA) with the low two bits dropped is 4 floor(u nA)), one LOP3 that
// joins it with the row field of the previous table entry, one shared-memory load of the new entry, two LOP3 on
// the rock bits (good? / the rock is bad afterwards), the reward address, its load, and three fp64 operations;
// the live-step counter adds the entry's top bit.  A trajectory that left the grid sits in an absorbing row whose
// entries add +0.0, so there is no per-step "alive" select; the step limit and the discount underflow are the same
// for every lane and are handled by the loop bounds.
// NT trajectories per lane (1 or 2), interleaved (rollouts_per_leaf > 32 on a single warp).  Both start
// at step 0, so they share the discount sequence; each accumulates exactly as it would alone.
// WIDE: more than 16 actions (k > 11 rocks); actions from 15 on are clamped into column 15.
template <bool CHECK_DISC, int NT, bool WIDE>
__device__ __forceinline__ void rollout_rocksample_impl(const DevModel& m, const RsTab* tb, uint32_t s, long long steps,
                                                        uint32_t k0, uint32_t k1, const uint32_t (&tagword)[NT],
                                                        const bool (&enabled)[NT], uint32_t q, uint32_t epoch,
                                                        int& nsteps, double (&total)[NT]) {
  const double gamma = m.gamma;
  const uint32_t nA4 = 4u * (uint32_t)m.nA;
  const int nstep_max = (int)(steps < 0x7fffffffll ? steps : 0x7fffffffll);
  const unsigned char* mvb = (const unsigned char*)tb->mv;
  const unsigned char* rtb = (const unsigned char*)tb->rtab2;
  uint32_t ent[NT];    // the table entry of the last step: its row field is where the trajectory is now
  uint32_t rocks[NT];  // bit 14 + i: rock i is good
  uint32_t step[NT];
  Philox4 p[NT];
  double disc = 1.0;
  bool any = false;
#pragma unroll
  for (int t = 0; t < NT; ++t) {
    const bool live = enabled[t] && !(s & POMCP_RS_TERMINAL) && nstep_max > 0;
    ent[t] = live ? (s & 0xffu) * 64u : RS_DEAD_ROW;
    rocks[t] = ((s >> 8) & 0xffffu) << RS_ROCK_SHIFT;
    total[t] = 0.0;
    step[t] = 0;
    any = any || live;
    p[t] = philox4x32_10(0u, tagword[t], q, epoch, k0, k1);
  }
  // one step of trajectory t with the 16-bit draw j of its current block
  auto one_step = [&](int t, int j) {
    // rand(rng, actions) = floor(u * nA) with a 16-bit u: a Philox block feeds 8 steps (low half of a
    // word first).  P(a) is within nA / 2^16 (relative) of 1 / nA.
    const uint32_t wj = p[t].w[j >> 1];
    uint32_t a4 = __umulhi((j & 1) ? (wj & 0xffff0000u) : (wj << 16), nA4);  // 4 a + two junk bits
    if (WIDE) a4 = min(a4, 63u);
    uint32_t off;  // (ent & RS_ROW_MASK) | (a4 & ~RS_ROW_MASK) as ONE LOP3 (the compiler splits the two masks)
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(off) : "r"(ent[t]), "r"(a4), "n"(RS_ROW_MASK));
    const uint32_t e = *(const uint32_t*)(mvb + off);
    step[t] += e >> 31;
    // sample on a rock: +good / -bad, and the rock is bad afterwards either way
    const uint32_t good = rocks[t] & e;
    rocks[t] &= ~e;
    total[t] += disc * *(const double*)(rtb + (e & 0x30u) + (good ? 8u : 0u));  // a finished trajectory adds +0.0
    ent[t] = e;
  };
  int base = 0;  // steps done by the lanes still on the grid
  for (uint32_t blk = 0; any; ++blk) {
    // The next block's Philox rounds are written INSIDE each arm: they have to sit in the same basic block as the
    // steps to be scheduled next to them (in front of the branch they ran first, and the look-up chain started 110
    // cycles late in every block).
    Philox4 pn[NT];
    bool stop = false;  // uniform: step limit reached or discount underflow
    if (!CHECK_DISC && base + 8 <= nstep_max) {
#pragma unroll
      for (int t = 0; t < NT; ++t) pn[t] = philox4x32_10(blk + 1u, tagword[t], q, epoch, k0, k1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
#pragma unroll
        for (int t = 0; t < NT; ++t) one_step(t, j);
        disc *= gamma;
      }
      base += 8;
      stop = base >= nstep_max;
    } else {
#pragma unroll
      for (int t = 0; t < NT; ++t) pn[t] = philox4x32_10(blk + 1u, tagword[t], q, epoch, k0, k1);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        if (!stop) {
#pragma unroll
          for (int t = 0; t < NT; ++t) one_step(t, j);
          disc *= gamma;
          ++base;
          stop = base >= nstep_max || (CHECK_DISC && !(disc > 0.0));
        }
      }
    }
    any = false;
#pragma unroll
    for (int t = 0; t < NT; ++t) {
      p[t] = pn[t];
      any = any || (!stop && (ent[t] & 0x3fc0u) != RS_DEAD_ROW);
    }
  }
  nsteps = 0;
#pragma unroll
  for (int t = 0; t < NT; ++t) nsteps += (int)step[t];
}
template <bool CHECK_DISC, int NT>
__device__ __forceinline__ void rollout_rocksample(const DevModel& m, const RsTab* tb, uint32_t s, long long steps,
                                                   uint32_t k0, uint32_t k1, const uint32_t (&tagword)[NT],
                                                   const bool (&enabled)[NT], uint32_t q, uint32_t epoch, int& nsteps,
                                                   double (&total)[NT]) {
  if (m.nA > 16) rollout_rocksample_impl<CHECK_DISC, NT, true>(m, tb, s, steps, k0, k1, tagword, enabled, q, epoch, nsteps, total);
  else rollout_rocksample_impl<CHECK_DISC, NT, false>(m, tb, s, steps, k0, k1, tagword, enabled, q, epoch, nsteps, total);
}

template <bool CHECK_DISC>
__device__ __forceinline__ double rollout_rocksample(const DevModel& m, const RsTab* tb, uint32_t s, long long steps,
                                                  uint32_t k0, uint32_t k1, uint32_t tagword, uint32_t q,
                                                  uint32_t epoch, int& nsteps) {
  const uint32_t tw[1] = {tagword};
  const bool en[1] = {true};
  double total[1];
  rollout_rocksample<CHECK_DISC, 1>(m, tb, s, steps, k0, k1, tw, en, q, epoch, nsteps, total);
  return total[0];
}

// ---- K3: RolloutSimulator loop under the uniform-random policy (src/rollout.jl:77-84;
// loop body = POMDPToolbox.RolloutSimulator).  One thread = one trajectory,
// state in registers, one Philox block feeds 4/DPS steps; the observation is never generated
// (VoidUpdater discards it, src/rollout.jl:104).
template <int PID, bool CHECK_DISC = true>
__device__ __forceinline__ double rollout_random(const DevModel& m, const RsTab* tb, typename ModelT<PID>::State s,
                                                 long long steps, uint32_t k0, uint32_t k1, uint32_t tagword,
                                                 uint32_t q, uint32_t epoch, int& nsteps) {
  if constexpr (PID == POMCP_ROCKSAMPLE)
    return rollout_rocksample<CHECK_DISC>(m, tb, s, steps, k0, k1, tagword, q, epoch, nsteps);
  else
    return rollout_generic<PID, CHECK_DISC>(m, s, steps, k0, k1, tagword, q, epoch, nsteps);
}

// Baby only: rollout under FeedWhenCrying (feed iff the previous observation was crying), starting
// from the leaf node's observation label (src/rollout.jl:105-106).  Words per step: transition, observation.
__device__ __forceinline__ double rollout_fwc(const DevModel& m, uint32_t s, long long steps, uint64_t last_obs,
                                              uint32_t k0, uint32_t k1, uint32_t tagword, uint32_t q, uint32_t epoch,
                                              int& nsteps) {
  using M = ModelT<POMCP_BABY>;
  const double gamma = m.gamma;
  double disc = 1.0, total = 0.0;
  long long step = 0;
  bool live = steps > 0;
  for (uint32_t blk = 0; live; ++blk) {
    Philox4 p = philox4x32_10(blk, tagword, q, epoch, k0, k1);
#pragma unroll
    for (int j = 0; j < 2; ++j) {
      if (live) {
        int a = last_obs ? 1 : 0;
        uint32_t sp;
        uint64_t o;
        double r;
        M::template step<true, double>(m, s, a, u01(p.w[2 * j]), u01(p.w[2 * j + 1]), 0.0, sp, o, r);
        total += disc * r;
        s = sp;
        last_obs = o;
        disc *= gamma;
        ++step;
        live = disc > 0.0 && step < steps;
      }
    }
  }
  nsteps = (int)step;
  return total;
}

// Baby only: FORollout(FeedWhenCrying()) (src/rollout.jl:86-92): the policy sees the state, feeds iff hungry.
// One Philox word per step (transition), no observation.
__device__ __forceinline__ double rollout_fo_fwc(const DevModel& m, uint32_t s, long long steps, uint32_t k0,
                                                 uint32_t k1, uint32_t tagword, uint32_t q, uint32_t epoch,
                                                 int& nsteps) {
  using M = ModelT<POMCP_BABY>;
  const double gamma = m.gamma;
  double disc = 1.0, total = 0.0;
  long long step = 0;
  bool live = steps > 0;
  for (uint32_t blk = 0; live; ++blk) {
    Philox4 p = philox4x32_10(blk, tagword, q, epoch, k0, k1);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (live) {
        uint32_t sp;
        uint64_t o;
        double r;
        M::template step<false, double>(m, s, s ? 1 : 0, u01(p.w[j]), 0.0, 0.0, sp, o, r);
        total += disc * r;
        s = sp;
        disc *= gamma;
        ++step;
        live = disc > 0.0 && step < steps;
      }
    }
  }
  nsteps = (int)step;
  return total;
}

__device__ __forceinline__ long long state_index(uint32_t s) { return (long long)(s & 0xffffffu); }
__device__ __forceinline__ long long state_index(const LDState&) { return -1; }

// Leaf evaluation by one warp: lane l runs rollout r_base + l (rollout team, batched mode) or, on a
// single warp with R > 32 (exact mode), rollouts l and l + 32; the per-lane sums are added by the fixed
// butterfly, so the oracle reproduces the mean bit for bit.
template <int PID>
__device__ __forceinline__ double leaf_rollouts(const DevModel& m, const RsTab* tb, const SearchCfg& cfg,
                                                const RngKey& rk, typename ModelT<PID>::State s, long long steps,
                                                uint64_t last_o, uint32_t q, int lane, int& nsteps_lane, int r_base = 0);

// ---- child lookup / insert (src/solver.jl:132-142).  Lane 0 only.  A node id that loses an insertion
// race here is not recycled (it stays an unreachable, zero-initialised slot whose parent link is 0, which
// the pool compaction of reroot.cuh reads as "not in the tree").
// SKIP_LOAD: the caller has just read the child slot as empty (direct table only), go straight to the insert.
// `created` tells the caller that it published the node itself (then it is, like in the reference,
// the natural first visitor of that node).
// `claim` (direct table only): publish the id with CHILD_EXPANDED; the returned value keeps that bit.
template <int PID, bool SKIP_LOAD = false>
__device__ __forceinline__ uint32_t find_or_insert_child(const Pool& P, const DevModel& m, uint32_t slot,
                                                         uint64_t obs, bool& created, bool claim = false,
                                                         double bel = 0.0) {
  using M = ModelT<PID>;
  created = false;
  if (!M::HASHED) {
    uint32_t* cs = &P.child[(size_t)slot * m.nO + (uint32_t)obs];
    if (!SKIP_LOAD) {
      uint32_t c = __ldcg(cs);
      if (c != 0) return c;
    }
    uint32_t fresh = (uint32_t)atomicAdd(&P.ctr->node_count, 1ull);
    if (fresh >= P.max_nodes) {
      P.ctr->pool_overflow = 1;
      return 0;
    }
    P.node_pslot[fresh] = slot;  // parent link, used by the pool compaction (reroot.cuh)
    if (P.node_bel) { P.node_bel[fresh] = bel; __threadfence(); }  // new_belief = update(node_belief_updater, h.B, a, o)
    const uint32_t pub = claim ? (fresh | CHILD_EXPANDED) : fresh;
    uint32_t old = atomicCAS(cs, 0u, pub);
    if (old == 0) { created = true; return pub; }
    P.node_pslot[fresh] = 0;  // lost the race: leave the slot detached
    return old;
  } else {
    uint64_t h = mix64((uint64_t)slot * 0x9E3779B97F4A7C15ull ^ obs) & P.ht_mask;
    uint32_t fresh = 0;  // allocated at the first empty slot; kept while probing past other keys' insertions
    for (;;) {
      uint32_t id = __ldcg(&P.ht[h]);
      if (id == 0) {
        if (fresh == 0) {
          fresh = (uint32_t)atomicAdd(&P.ctr->node_count, 1ull);
          if (fresh >= P.max_nodes) {
            P.ctr->pool_overflow = 1;
            return 0;
          }
          P.node_pslot[fresh] = slot;
          P.node_obs[fresh] = obs;
          __threadfence();
        }
        uint32_t old = atomicCAS(&P.ht[h], 0u, fresh);
        if (old == 0) { created = true; return fresh; }
        id = old;  // another insertion took this slot: same key -> use it, other key -> keep probing
      }
      if (__ldcg(&P.node_pslot[id]) == slot && __ldcg(&P.node_obs[id]) == obs) {
        if (fresh != 0) P.node_pslot[fresh] = 0;  // lost the race for this very key: leave the slot detached
        return id;
      }
      h = (h + 1) & P.ht_mask;
    }
  }
}

// Value-returning atomic add whose result is NOT needed right away.  `atomicAdd` under `if (lane == 0)` is
// compiled into the warp-aggregated form (vote, one ATOMG, shuffle of the result to the lanes), and that
// shuffle waits for the round trip on the spot; the plain PTX instruction leaves the wait to the first use.
__device__ __forceinline__ unsigned long long atom_add_u64_async(unsigned long long* addr, unsigned long long v,
                                                                  unsigned any_runtime_value) {
  // ptxas aggregates the PTX instruction as well when it can prove the address warp-uniform.  The callers are
  // lane 0, so laneid * (anything it cannot fold) is an offset of zero that keeps the address lane-dependent.
  unsigned laneid;
  asm volatile("mov.u32 %0, %%laneid;" : "=r"(laneid));
  const unsigned long long off = (unsigned long long)laneid * (unsigned long long)any_runtime_value * 8ull;
  unsigned long long r;
  asm volatile("atom.global.add.u64 %0, [%1], %2;" : "=l"(r) : "l"(__cvta_generic_to_global(addr) + off), "l"(v) : "memory");
  return r;
}

// Batched workers, direct table: the same insertion with a node id the worker allocated ahead of time (`spare`,
// 0 = none yet), so the critical path holds one round trip (the CAS) instead of two, and an id that loses
// the race is kept for the next insertion instead of being dropped.  The replacement id is requested right
// after a successful insertion and is not waited for until the next one.  Ids then no longer grow along every
// root-to-leaf path (a spare can be older than the parent it ends up under); the pool compaction walks
// parent links to the root instead of relying on that order.
__device__ __forceinline__ uint32_t insert_child_spare(const Pool& P, int nO, uint32_t slot, uint64_t obs, bool& created,
                                                       bool claim, uint32_t& spare, double bel = 0.0) {
  created = false;
  uint32_t* cs = &P.child[(size_t)slot * nO + (uint32_t)obs];
  uint32_t fresh = spare;
  if (fresh == 0) fresh = (uint32_t)atomicAdd(&P.ctr->node_count, 1ull);
  if (fresh >= P.max_nodes) {
    P.ctr->pool_overflow = 1;
    return 0;
  }
  P.node_pslot[fresh] = slot;  // parent link, used by the pool compaction (reroot.cuh)
  if (P.node_bel) { P.node_bel[fresh] = bel; __threadfence(); }  // the node's belief is in place before its id is visible
  const uint32_t pub = claim ? (fresh | CHILD_EXPANDED) : fresh;
  const uint32_t old = atomicCAS(cs, 0u, pub);
  if (old == 0) {
    created = true;
    spare = (uint32_t)atom_add_u64_async(&P.ctr->node_count, 1ull, P.max_nodes);
    return pub;
  }
  P.node_pslot[fresh] = 0;  // lost the race: the id stays detached and is offered again next time
  spare = fresh;
  return old;
}

// ---- observation progressive widening (src/solver.jl:223-262).  Every generated (child, state) pair of an
// action node is an event e = 0, 1, ...; a widened step re-uses event floor(u * #events): the child with
// probability N(child) / sum N and one of the states pushed into it, uniformly — what sample(os, WeightVec(N))
// followed by rand(hao.B) draws.
__device__ __forceinline__ bool dpw_generates(const SearchCfg& cfg, uint32_t act_n, uint32_t n_children) {
  // length(best_node.children) <= k_observation * best_node.N^alpha_observation
  const double lim = act_n == 0 ? 0.0
                   : cfg.alpha_obs == 0.5 ? cfg.k_obs * sqrt((double)act_n)
                   : cfg.k_obs * det_exp(cfg.alpha_obs * det_log((double)act_n));
  return (double)n_children <= lim;
}
// length(h.children) <= k_action * total_N^alpha_action (src/solver.jl:173)
__device__ __forceinline__ bool dpw_widens_actions(const SearchCfg& cfg, unsigned long long total_n, int n_children) {
  const double lim = total_n == 0 ? 0.0
                   : cfg.alpha_act == 0.5 ? cfg.k_act * sqrt((double)total_n)
                   : cfg.k_act * det_exp(cfg.alpha_act * det_log((double)total_n));
  return (double)n_children <= lim;
}
template <class State>
__device__ __forceinline__ bool ev_insert(const Pool& P, uint32_t slot, uint32_t e, uint32_t node, const State& sp) {
  const unsigned long long key = (((unsigned long long)slot << 32) | e) + 1ull;
  unsigned long long h = mix64(key) & P.ev_mask;
  for (unsigned probe = 0; probe < 4096; ++probe) {
    const unsigned long long old = atomicCAS(&P.ev_key[h], 0ull, ~0ull);  // claim, fill, then publish the key
    if (old == 0ull) {
      P.ev_node[h] = node;
      ((State*)P.ev_state)[h] = sp;
      __threadfence();
      atomicExch(&P.ev_key[h], key);
      return true;
    }
    h = (h + 1) & P.ev_mask;
  }
  P.ctr->pool_overflow = 1;
  return false;
}
template <class State>
__device__ __forceinline__ bool ev_lookup(const Pool& P, uint32_t slot, uint32_t e, uint32_t& node, State& sp) {
  const unsigned long long key = (((unsigned long long)slot << 32) | e) + 1ull;
  unsigned long long h = mix64(key) & P.ev_mask;
  for (unsigned probe = 0; probe < 4096; ++probe) {
    const unsigned long long k = ld_volatile_u64g(&P.ev_key[h]);
    if (k == key) {
      __threadfence();
      node = __ldcg(&P.ev_node[h]);
      sp = ((const State*)P.ev_state)[h];
      return true;
    }
    if (k == 0ull) return false;  // not (yet) published
    h = (h + 1) & P.ev_mask;
  }
  return false;
}

// read-only lookup used by update / inspection
template <int PID>
__device__ __forceinline__ uint32_t find_child(const Pool& P, const DevModel& m, uint32_t slot, uint64_t obs) {
  using M = ModelT<PID>;
  if (!M::HASHED) {
    if (obs >= (uint64_t)m.nO) return 0;
    return __ldcg(&P.child[(size_t)slot * m.nO + (uint32_t)obs]) & CHILD_ID_MASK;
  } else {
    uint64_t h = mix64((uint64_t)slot * 0x9E3779B97F4A7C15ull ^ obs) & P.ht_mask;
    for (;;) {
      uint32_t id = __ldcg(&P.ht[h]);
      if (id == 0) return 0;
      if (__ldcg(&P.node_pslot[id]) == slot && __ldcg(&P.node_obs[id]) == obs) return id;
      h = (h + 1) & P.ht_mask;
    }
  }
}

__device__ __forceinline__ uint32_t two_state(uint32_t, bool one) { return one ? 1u : 0u; }
__device__ __forceinline__ LDState two_state(const LDState& s, bool) { return s; }  // never called for LightDark
__device__ __forceinline__ uint32_t baby_state(uint32_t s) { return s; }
__device__ __forceinline__ uint32_t baby_state(const LDState&) { return 0u; }  // never called for LightDark

__device__ __forceinline__ uint32_t shfl_state(uint32_t v, int src) { return __shfl_sync(0xffffffffu, v, src); }
__device__ __forceinline__ LDState shfl_state(const LDState& v, int src) {
  LDState o;
  o.status = __shfl_sync(0xffffffffu, v.status, src);
  o.y = __shfl_sync(0xffffffffu, v.y, src);
  return o;
}

// fixed-shape butterfly sum over the warp: the oracle adds in the same binary tree
__device__ __forceinline__ double warp_tree_sum(double x) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) x = x + __shfl_xor_sync(0xffffffffu, x, off);
  return x;
}

template <int PID>
__device__ __forceinline__ double leaf_rollouts(const DevModel& m, const RsTab* tb, const SearchCfg& cfg,
                                                const RngKey& rk, typename ModelT<PID>::State s, long long steps,
                                                uint64_t last_o, uint32_t q, int lane, int& nsteps_lane, int r_base) {
  const int R = cfg.rollouts;
  const uint32_t tag0 = TAG_ROLLOUT | rk.tagrank;
  double ret = 0.0;
  nsteps_lane = 0;
  if (R > 32 && cfg.team_warps <= 1) {  // one warp, two trajectories per lane (warp-uniform)
    if constexpr (PID == POMCP_ROCKSAMPLE) {
      const uint32_t tw[2] = {tag0 | ((uint32_t)lane << 16), tag0 | ((uint32_t)(lane + 32) << 16)};
      const bool en[2] = {true, lane + 32 < R};
      double tot[2];
      rollout_rocksample<false, 2>(m, tb, s, steps, rk.key0, rk.key1, tw, en, q, rk.epoch, nsteps_lane, tot);
      ret = tot[0] + tot[1];
    } else {
      double tot[2] = {0.0, 0.0};
#pragma unroll
      for (int t = 0; t < 2; ++t) {
        const int r = lane + 32 * t;
        int ns = 0;
        if (r < R) {
          const uint32_t tagword = tag0 | ((uint32_t)r << 16);
          if (PID == POMCP_BABY && cfg.estimator == POMCP_EST_FWC_ROLLOUT)
            tot[t] = rollout_fwc(m, (uint32_t)baby_state(s), steps, last_o, rk.key0, rk.key1, tagword, q, rk.epoch, ns);
          else if (PID == POMCP_BABY && cfg.estimator == POMCP_EST_FO_FWC_ROLLOUT)
            tot[t] = rollout_fo_fwc(m, (uint32_t)baby_state(s), steps, rk.key0, rk.key1, tagword, q, rk.epoch, ns);
          else
            tot[t] = rollout_random<PID, false>(m, tb, s, steps, rk.key0, rk.key1, tagword, q, rk.epoch, ns);
        }
        nsteps_lane += ns;
      }
      ret = tot[0] + tot[1];
    }
  } else if (r_base + lane < R) {  // rollout team: warp w of the worker runs rollouts 32 w .. 32 w + 31
    const uint32_t tagword = tag0 | ((uint32_t)(r_base + lane) << 16);
    if (PID == POMCP_BABY && cfg.estimator == POMCP_EST_FWC_ROLLOUT)
      ret = rollout_fwc(m, (uint32_t)baby_state(s), steps, last_o, rk.key0, rk.key1, tagword, q, rk.epoch, nsteps_lane);
    else if (PID == POMCP_BABY && cfg.estimator == POMCP_EST_FO_FWC_ROLLOUT)
      ret = rollout_fo_fwc(m, (uint32_t)baby_state(s), steps, rk.key0, rk.key1, tagword, q, rk.epoch, nsteps_lane);
    else
      ret = rollout_random<PID, false>(m, tb, s, steps, rk.key0, rk.key1, tagword, q, rk.epoch, nsteps_lane);
  }
  return ret;
}

// ---- one simulation, run by one warp: simulate() of src/solver.jl:78-157.
//  * K1 root draw (src/solver.jl:48-49);
//  * K2 descent: control flow is warp-uniform, lanes hold the actions for the UCB arg-max, lane 0
//    issues the atomics.  Visits are counted on arrival: the value returned by the node_N ticket
//    is exactly the reference's h.N in a sequential run, and with many simulations in flight it
//    lets each one see the others' unfinished visits, so they spread over actions like sequential
//    UCB would instead of piling up on one arm;
//  * K3 leaf evaluation: lane r runs rollout r (R <= 32), returns are reduced by shuffles;
//  * K4 backup bottom-up (src/solver.jl:144-156) and particle pushes.
// Returns false when the query was consumed by a terminal root sample.
template <int PID, bool EXACT, bool REPLAY>
__device__ __forceinline__ bool simulate_one(const Pool& P, const DevModel& m, const RsTab* tb, const SearchCfg& cfg,
                                             const PathBuf& pb, int slot_idx, uint32_t q, uint32_t root_ticket,
                                             ReplayCursor& rc, uint32_t& spare, LocalCtr& lc) {
  using M = ModelT<PID>;
  using State = typename M::State;
  const int lane = threadIdx.x & 31;
  const int nA = m.nA;
  State* path_state = (State*)pb.path_state;

  // rand(rng, b): src/solver.jl:48 -> src/updater.jl:53-55
  double u[4];
  get_draws<REPLAY>(cfg, rc, TAG_ROOT, 0, q, u);
  if (REPLAY) rc.ucur -= 2;  // the root draw consumes two uniforms
  State s;
  if (cfg.bel != nullptr) {
    long long idx = (long long)(u[0] * (double)cfg.n_bel);
    s = ((const State*)cfg.bel)[idx];
  } else if (M::TWO_STATE && cfg.exact2) {
    s = two_state(State{}, u[0] < __ldcg(&P.node_bel[cfg.root]));  // rand(rng, BoolDistribution(p))
  } else {
    s = M::initial(m, u[0], u[1]);
  }
  if (M::terminal(m, s)) {  // src/solver.jl:49: the query is consumed
    lc.skips++;
    return false;
  }

  uint32_t h = cfg.root;
  uint32_t hN = root_ticket;
  uint64_t last_o = 0;  // observation label of h (the leaf's label seeds the FeedWhenCrying rollout)
  int depth = 0;
  bool want_rollout = false, aborted = false;
  // Two dependent L2 round trips per level: (1) the node's action statistics together with its
  // child ids, (2) the arrival ticket of the chosen child together with that child's expansion flag.
  uint32_t fl = 0;  // lane 0 only: expansion flag of h
  if (lane == 0) fl = __ldcg(&P.node_flags[h]);
  for (;;) {
    // src/solver.jl:82-86
    if (depth >= cfg.levels || M::terminal(m, s)) break;
    // src/solver.jl:87-107: first non-cut-off visitor expands and evaluates the leaf.
    // The flag races with other workers: lane 0 reads (and claims) it and broadcasts the outcome.
    // A per-lane load could return different values to different lanes of this warp (lanes are
    // not guaranteed to execute the load together) and split the warp across shuffle sites.
    int i_expand = 0;
    uint32_t amask = 0xffffffffu;  // existing ActNodes (all of them unless actions are widened progressively)
    unsigned long long totalN = hN;
    if (cfg.dpa) {
      // src/solver.jl:172-186 (exact mode: sequential, plain loads / stores)
      amask = __ldcg(&P.node_amask[h]);
      const uint32_t nl = (lane < nA && ((amask >> lane) & 1u)) ? __ldcg(&P.act_N[h * (uint32_t)nA + (uint32_t)lane]) : 0u;
      totalN = __reduce_add_sync(0xffffffffu, nl);
      if (dpw_widens_actions(cfg, totalN, __popc(amask))) {
        double ua[4];
        get_draws<REPLAY>(cfg, rc, TAG_TREE, (uint32_t)depth | 0x80000000u, q, ua);
        if (REPLAY) rc.ucur -= 3;  // next_action consumes one uniform
        const uint32_t bit = 1u << (int)(ua[0] * (double)nA);
        if (!(amask & bit)) {
          amask |= bit;
          if (lane == 0) P.node_amask[h] = amask;
        }
        i_expand = __popc(amask) <= 1;
      }
    } else {
      if (lane == 0 && !(fl & 1u)) {
        uint32_t old;
        if (EXACT) { old = fl; P.node_flags[h] = fl | 1u; }
        else old = atomicOr(&P.node_flags[h], 1u);
        i_expand = !(old & 1u);
      }
      i_expand = __shfl_sync(0xffffffffu, i_expand, 0);
    }
    if (i_expand) {
      want_rollout = depth > 0;
      break;
    }
    // UCB1: src/solver.jl:109-124 (child ids of every action are fetched by the same lanes)
    double crit = 0.0;
    bool valid = false;
    uint32_t kid[M::NO > 0 ? M::NO : 1];
#pragma unroll
    for (int k = 0; k < (M::NO > 0 ? M::NO : 1); ++k) kid[k] = 0;
    if (lane < nA) {
      uint32_t sl = h * (uint32_t)nA + (uint32_t)lane;
      uint32_t n = __ldcg(&P.act_N[sl]);
      double v = __ldcg(&P.act_V[sl]);
      uint32_t mm = 0;
      if (!EXACT) mm = __ldcg(&P.act_M[sl]);
      if (!M::HASHED) {
#pragma unroll
        for (int k = 0; k < M::NO; ++k) kid[k] = __ldcg(&P.child[(size_t)sl * M::NO + k]);
      }
      if (!EXACT) v = (mm > 0 && fabs(cfg.init_V) != INFINITY) ? v / (double)mm : cfg.init_V;
      unsigned long long hn = cfg.dpa ? totalN : (EXACT ? hN : (hN < 1u ? 1u : hN));
      if (n == 0 && (cfg.dpw_solver ? hn <= 1 : hn == 1)) crit = v;  // src/solver.jl:112 / :210
      else if (n == 0 && v == -INFINITY) crit = INFINITY;
      else crit = v + cfg.c * sqrt(det_log((double)hn) / (double)n);
      valid = (crit == crit) && ((amask >> lane) & 1u);
    }
    int best_a = lane;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      double oc = __shfl_xor_sync(0xffffffffu, crit, off);
      int oa = __shfl_xor_sync(0xffffffffu, best_a, off);
      bool ov = __shfl_xor_sync(0xffffffffu, (int)valid, off) != 0;
      bool take = ov && (!valid || oc > crit || (oc == crit && oa > best_a));  // ">=": last index wins ties
      if (take) { crit = oc; best_a = oa; valid = true; }
    }
    valid = __shfl_sync(0xffffffffu, (int)valid, 0) != 0;
    best_a = __shfl_sync(0xffffffffu, best_a, 0);
    if (!valid) best_a = nA - 1;
    const uint32_t slot = h * (uint32_t)nA + (uint32_t)best_a;

    // src/solver.jl:126
    get_draws<REPLAY>(cfg, rc, TAG_TREE, (uint32_t)depth, q, u);
    State sp;
    uint64_t o;
    double r;
    M::template step<true, double>(m, s, best_a, u[0], u[2], u[3], sp, o, r);

    // src/solver.jl:132-142 + arrival counting
    uint32_t pre = 0;  // prefetched child id (ids never change once published; 0 = not there yet)
    if (!M::HASHED) {
      uint32_t mine = kid[0];
#pragma unroll
      for (int k = 1; k < M::NO; ++k) mine = (o == (uint64_t)k) ? kid[k] : mine;
      pre = __shfl_sync(0xffffffffu, mine, best_a);
    }
    uint32_t hao = 0, ticket = 0;
    if (cfg.dpw) {
      // POMCPDPWSolver, enable_action_pw = false (exact mode only reaches here sequentially)
      int reuse = 0;
      if (lane == 0) {
        const uint32_t bn = __ldcg(&P.act_N[slot]);  // best_node.N before this visit
        const uint32_t nc = __ldcg(&P.act_nc[slot]), ng = __ldcg(&P.act_ng[slot]);
        P.act_N[slot] = bn + 1u;
        if (dpw_generates(cfg, bn, nc)) {
          bool created_ = false;
          hao = find_or_insert_child<PID>(P, m, slot, o, created_);
          if (hao != 0) {
            if (created_) P.act_nc[slot] = nc + 1u;
            P.act_ng[slot] = ng + 1u;
            if (!ev_insert<State>(P, slot, ng, hao, sp)) hao = 0;
          }
          if (hao != 0) {
            ticket = __ldcg(&P.node_N[hao]) + 1u;  // hao.N += 1 before the recursive call (src/solver.jl:257-262)
            P.node_N[hao] = ticket;
          }
        } else {
          const uint32_t e = (uint32_t)(u[1] * (double)ng);  // word 1 of the level's block: unused by every model
          reuse = 1;
          if (!ev_lookup<State>(P, slot, e, hao, sp)) hao = 0;
          if (hao != 0) ticket = __ldcg(&P.node_N[hao]);
        }
        if (hao != 0) {
          fl = __ldcg(&P.node_flags[hao]);
          size_t pi = (size_t)depth * pb.stride + slot_idx;
          pb.path_slot[pi] = slot;
          pb.path_node[pi] = hao;
          pb.path_r[pi] = r;
          path_state[pi] = sp;
        }
      }
      reuse = __shfl_sync(0xffffffffu, reuse, 0);
      if (reuse) sp = shfl_state(sp, 0);
    } else if (lane == 0) {
      if (EXACT) P.act_N[slot] = __ldcg(&P.act_N[slot]) + 1u; else atomicAdd(&P.act_N[slot], 1u);
      bool created_;
      if (pre) hao = pre;
      else hao = find_or_insert_child<PID>(P, m, slot, o, created_, false,
                                           (M::TWO_STATE && cfg.exact2) ? bayes2<PID>(m, P.node_bel[h], best_a, o) : 0.0);
      if (hao != 0) {
        if (EXACT) { ticket = __ldcg(&P.node_N[hao]); P.node_N[hao] = ticket + 1u; }
        else ticket = atomicAdd(&P.node_N[hao], 1u);
        fl = __ldcg(&P.node_flags[hao]);  // independent of the ticket: both are in flight together
        size_t pi = (size_t)depth * pb.stride + slot_idx;
        pb.path_slot[pi] = slot;
        pb.path_node[pi] = hao;
        pb.path_r[pi] = r;
        path_state[pi] = sp;
      }
    }
    hao = __shfl_sync(0xffffffffu, hao, 0);
    ticket = __shfl_sync(0xffffffffu, ticket, 0);
    if (hao == 0) { aborted = true; break; }
    h = hao;
    hN = ticket;
    s = sp;
    last_o = o;
    ++depth;
  }
  if (aborted) return true;  // pool exhausted: the search reports POOL_EXHAUSTED

  // leaf evaluation, src/solver.jl:97-103
  double leaf = 0.0;
  if (want_rollout) {
    long long steps_to_eps = (long long)ceil(cfg.log_ratio - (double)depth);
    long long steps = min(cfg.max_depth - (long long)depth, steps_to_eps);
    if (cfg.dpw_solver) steps = depth;  // estimate_value(..., s, h, depth), src/solver.jl:182,199
    double est;
    if (REPLAY) {
      est = (rc.rcur < cfg.n_rets) ? cfg.rets[rc.rcur] : 0.0;
      rc.rcur++;
      lc.rollouts++;
    } else if (cfg.estimator == POMCP_EST_CONSTANT) {
      est = cfg.est_value;  // src/rollout.jl:10
      lc.rollouts++;
    } else if (cfg.estimator == POMCP_EST_STATE_VALUE) {  // value(policy, s), src/rollout.jl:67-69
      const long long si = state_index(s);
      est = (M::terminal(m, s) || si < 0 || si >= cfg.n_state_values) ? 0.0 : cfg.state_values[si];
      lc.rollouts++;
    } else {
      int nsteps = 0;
      const RngKey rk{cfg.key0, cfg.key1, cfg.rank << 8, cfg.epoch};
      const double ret = leaf_rollouts<PID>(m, tb, cfg, rk, s, steps, last_o, q, lane, nsteps);
      est = warp_tree_sum(ret) / (double)cfg.rollouts;
      nsteps = __reduce_add_sync(0xffffffffu, nsteps);
      lc.rollouts += cfg.rollouts;
      lc.rollout_steps += nsteps;
    }
    leaf = cfg.gpow[depth] * est;  // src/solver.jl:103 (the reference's extra gamma^depth is kept)
  }

  // particle log reservation (lane 0), then lanes write one level each
  unsigned long long log_base = ~0ull;
  if (cfg.belief_unweighted && depth > 0) {
    if (lane == 0) {
      if (EXACT) { log_base = P.ctr->log_count; P.ctr->log_count = log_base + depth; }
      else log_base = atomicAdd(&P.ctr->log_count, (unsigned long long)depth);
      if (log_base + depth > P.log_cap) { P.ctr->log_overflow = 1; log_base = ~0ull; }
    }
    log_base = __shfl_sync(0xffffffffu, log_base, 0);
  }
  __syncwarp();
  if (log_base != ~0ull) {
    State* log_state = (State*)P.log_state;
    for (int d = depth - 1 - lane; d >= 0; d -= 32) {
      // deepest level first, like the unwinding recursion (src/solver.jl:146-148)
      size_t pi = (size_t)d * pb.stride + slot_idx;
      unsigned long long li = log_base + (unsigned long long)(depth - 1 - d);
      P.log_node[li] = pb.path_node[pi];
      log_state[li] = path_state[pi];
    }
  }
  // backup, src/solver.jl:144-156
  if (lane == 0) {
    double R = leaf;
    for (int d = depth - 1; d >= 0; --d) {
      size_t pi = (size_t)d * pb.stride + slot_idx;
      R = pb.path_r[pi] + m.gamma * R;
      uint32_t slot = pb.path_slot[pi];
      if (EXACT) {
        double v = P.act_V[slot];
        if (v != -INFINITY) P.act_V[slot] = v + (R - v) / (double)P.act_N[slot];  // src/solver.jl:151-154
      } else {
        atomicAdd(&P.act_V[slot], R);
        atomicAdd(&P.act_M[slot], 1u);
      }
    }
    // b.N += 1 after simulate (src/solver.jl:51)
    if (EXACT) P.node_N[cfg.root] = __ldcg(&P.node_N[cfg.root]) + 1u;
    else { __threadfence(); atomicAdd(&P.node_N[cfg.root], 1u); }
  }
  lc.sims++;
  lc.tree_steps += depth;
  if ((unsigned long long)depth > lc.max_depth) lc.max_depth = depth;
  return true;
}

// ---------------------------------------------------------------------------------------------
// Batched mode.  Same algorithm as simulate_one, arranged for latency: a simulation is a chain of
// dependent L2/HBM accesses, and with a bounded number of simulations in flight the search rate
// is (simulations in flight) / (latency of one simulation).
//  * one memory round trip per level: the child id of every (action, observation) is fetched with
//    the node's statistics, so once the action and the observation are known the arrival ticket of
//    the child, its expansion flag and its statistics are all requested together (the statistics
//    speculatively: they are unused if this worker then expands the child);
//  * the recorded path lives in shared memory, the backup does not touch global memory for it;
//  * UCB1 in fp32 with the SFU (lg2/rcp/sqrt approx): the exploration term does not need 53 bits,
//    and the arg-max with the reference's tie rule is one REDUX.MAX on an order-preserving key plus
//    one ballot (highest lane among the maxima = "last wins on >=");
//  * the next query id is requested before the rollouts start, so that atomic is off the path.
// RockSample generative step for the tree (batched mode): same transition, reward and observation as
// ModelT<ROCKSAMPLE>::step<true>, on the shared-memory tables.  The sensor draw compares the Philox
// word with ceil(eta * 2^32) - 1, which is the same event as word * 2^-32 < eta.
__device__ __forceinline__ void rs_tree_step(const DevModel& m, const RsTab* tb, const uint32_t* eta_thr, uint32_t st,
                                             int a, uint32_t w_obs, uint32_t& sp, uint64_t& o, double& r) {
  // transition, event and sampled rock from the rollout table (one look-up instead of the move arithmetic, the
  // clamps and the rock map): entry layout at RsTab::mv
  const uint32_t e = tb->mv[(st & 0xffu) * 16u + (uint32_t)min(a, 15)];
  const uint32_t mask = ((e >> RS_ROCK_SHIFT) & 0xffffu) << 8;  // the sampled rock's bit in the state (0: none)
  const uint32_t good = (st & mask) ? 1u : 0u;
  r = *(const double*)((const unsigned char*)tb->rtab2 + (e & 0x30u) + good * 8u);
  uint32_t next = (st & 0xffffff00u & ~mask) | ((e >> 6) & 0xffu);
  next |= ((e & 0x30u) == 0x10u) ? POMCP_RS_TERMINAL : 0u;  // left the grid at the east edge
  if ((e & 0x30u) == 0x10u) next = (next & ~0xffu) | (st & 0xf0u) | (uint32_t)(m.rs_n - 1);  // x stays at the edge
  o = 2;  // none
  if (a >= 5) {
    const int rk = a - 5;
    const int x = (int)(st & 15u), y = (int)((st >> 4) & 15u);
    const uint32_t g = (st >> (8 + rk)) & 1u;
    const uint32_t correct = w_obs <= eta_thr[(y * m.rs_n + x) * m.rs_k + rk] ? 1u : 0u;
    o = (g == correct) ? 0 : 1;
  }
  sp = next;
}

// Lane-predicated memory operations: "lane 0 publishes" without splitting the warp (a branch around one
// instruction costs a BSSY / BSYNC pair, a re-convergence stall, and needs a __syncwarp() afterwards).
__device__ __forceinline__ void red_add_u32_if(bool p, uint32_t* addr) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p red.global.add.u32 [%1], 1;\n\t}" ::"r"((uint32_t)p),
               "l"(__cvta_generic_to_global(addr))
               : "memory");
}
__device__ __forceinline__ void red_or_u32_if(bool p, uint32_t* addr, uint32_t v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p red.global.or.b32 [%1], %2;\n\t}" ::"r"((uint32_t)p),
               "l"(__cvta_generic_to_global(addr)), "r"(v)
               : "memory");
}
__device__ __forceinline__ void st_u32_if(bool p, uint32_t* addr, uint32_t v) {  // generic address
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.u32 [%1], %2;\n\t}" ::"r"((uint32_t)p), "l"(addr), "r"(v)
               : "memory");
}
__device__ __forceinline__ void st_f64_if(bool p, double* addr, double v) {  // generic address
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %0, 0;\n\t@p st.f64 [%1], %2;\n\t}" ::"r"((uint32_t)p), "l"(addr), "d"(v)
               : "memory");
}

template <class State>
struct PathView {
  uint32_t* slot;
  uint32_t* node;
  double* r;
  State* state;
  int stride;
};

// .ftz: one MUFU each, no denormal pre/post-scaling (the arguments here are counts >= 1, their logarithms and ratios:
// never denormal, so the results are those of the non-ftz forms)
__device__ __forceinline__ float sfu_lg2(float x) { float y; asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sfu_rcp(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float sfu_sqrt(float x) { float y; asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
// order-preserving map float -> u32 (callers send NaN to 0)
__device__ __forceinline__ uint32_t f2ord(float x) {
  uint32_t b = __float_as_uint(x);
  return b ^ ((b >> 31) ? 0xffffffffu : 0x80000000u);
}

template <int PID>
struct ActStats {
  static constexpr int K = ModelT<PID>::NO > 0 ? ModelT<PID>::NO : 1;
  uint32_t n, mm;
  double w;
  uint32_t kid[K];
  uint32_t nc, ng;  // observation widening (hashed models): children / generation events of the action node
  uint32_t amask;   // action widening: existing ActNodes of the node (all ones otherwise)
  __device__ __forceinline__ void load(const Pool& P, uint32_t h, int nA, int lane, bool dpw = false, bool dpa = false) {
    n = 0; mm = 0; w = 0.0; nc = 0; ng = 0;
    amask = dpa ? __ldcg(&P.node_amask[h]) : 0xffffffffu;
#pragma unroll
    for (int k = 0; k < K; ++k) kid[k] = 0;
    if (lane < nA) {
      const uint32_t sl = h * (uint32_t)nA + (uint32_t)lane;
      n = __ldcg(&P.act_N[sl]);
      w = __ldcg(&P.act_V[sl]);
      mm = __ldcg(&P.act_M[sl]);
      if (dpw) {
        nc = __ldcg(&P.act_nc[sl]);
        ng = __ldcg(&P.act_ng[sl]);
      }
      if (!ModelT<PID>::HASHED) {
#pragma unroll
        for (int k = 0; k < ModelT<PID>::NO; ++k) kid[k] = __ldcg(&P.child[(size_t)sl * ModelT<PID>::NO + k]);
      }
    }
  }
};

// Rollout team: a worker is a CTA of `team_warps` warps.  Warp 0 walks the tree; at a leaf it posts the
// leaf state and every warp of the CTA runs 32 of the R = 32 * team_warps rollouts, so the leaf
// estimate averages up to 128 trajectories at the latency of 32 (the helpers sleep on a named
// barrier during the descent and take no issue slots).  A simulation's latency — and with it the
// staleness of the tree — does not grow with R.
template <class State>
struct LeafJob {
  State s;
  long long steps;
  unsigned long long last_o;
  uint32_t q;
  int cmd;  // 1: roll out, 0: the worker is done
  int bar;  // named barriers bar, bar + 1 belong to this team
  double partial[4];
  int psteps[4];
};
// Named barrier with an immediate id: with the id in a register ptxas reserves all 16 barriers for the CTA,
// and the SM's barrier pool then caps residency at four CTAs (the four-warp teams want seven).
template <int ID>
__device__ __forceinline__ void bar_imm(int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"n"(ID), "r"(nthreads) : "memory");
}
template <int NTEAMS>
__device__ __forceinline__ void team_bar(int id, int nthreads) {  // id in 1 .. 2 * NTEAMS (two per team)
  if (NTEAMS == 1) {
    if (id == 1) bar_imm<1>(nthreads); else bar_imm<2>(nthreads);
  } else if (NTEAMS == 2) {
    switch (id) {
      case 1: bar_imm<1>(nthreads); break;
      case 2: bar_imm<2>(nthreads); break;
      case 3: bar_imm<3>(nthreads); break;
      default: bar_imm<4>(nthreads); break;
    }
  } else {
    switch (id) {
      case 1: bar_imm<1>(nthreads); break;
      case 2: bar_imm<2>(nthreads); break;
      case 3: bar_imm<3>(nthreads); break;
      case 4: bar_imm<4>(nthreads); break;
      case 5: bar_imm<5>(nthreads); break;
      case 6: bar_imm<6>(nthreads); break;
      case 7: bar_imm<7>(nthreads); break;
      default: bar_imm<8>(nthreads); break;
    }
  }
}

// One simulation by one warp (batched mode).  `next_q` receives the worker's next query id
// (lane 0; requested early so its latency hides behind the rollouts).
template <int PID, bool DPA, int NTEAMS, bool DPO>
__device__ __forceinline__ void simulate_async(const Pool& P, const DevModel& m, const RsTab* tb,
                                               const uint32_t* eta_thr, const SearchCfg& cfg, const TreeView& tv,
                                               const PathView<typename ModelT<PID>::State>& pv,
                                               LeafJob<typename ModelT<PID>::State>* job, uint32_t q,
                                               LocalCtr& lc, unsigned long long& next_q, uint32_t& spare) {
  using M = ModelT<PID>;
  using State = typename M::State;
  const int lane = threadIdx.x & 31;
  const int nA = m.nA;
  const uint32_t tag_rank = tv.tagrank;
  const float c_f = (float)cfg.c;
  const bool v_finite = fabs(cfg.init_V) != INFINITY;
  const float initV_f = (float)cfg.init_V;

  // root statistics are requested first: they do not depend on the draw
  uint32_t h = tv.root;
  uint32_t hN = tv.root_N0 + q;
  ActStats<PID> st;
  st.load(P, h, nA, lane, DPO && cfg.dpw != 0, DPA);
  uint32_t fl = 0;
  if (lane == 0) fl = __ldcg(&P.node_flags[h]);
  __syncwarp();

  // rand(rng, b): src/solver.jl:48 -> src/updater.jl:53-55
  Philox4 pr = philox4x32_10(0u, TAG_ROOT | tag_rank, q, tv.epoch, tv.key0, tv.key1);
  const double u0 = u01(pr.w[0]), u1 = u01(pr.w[1]);
  State s;
  if (tv.bel != nullptr) s = ((const State*)tv.bel)[(long long)(u0 * (double)tv.n_bel)];
  else if (M::TWO_STATE && cfg.exact2) s = two_state(State{}, u0 < __ldcg(&P.node_bel[tv.root]));  // rand(rng, BoolDistribution(p))
  else s = M::initial(m, u0, u1);
  fl = __shfl_sync(0xffffffffu, fl, 0);
  if (M::terminal(m, s)) {  // src/solver.jl:49: the query is consumed
    lc.skips++;
    if (lane == 0) next_q = atomicAdd(&tv.ctl->next_query, 1ull);
    return;
  }

  uint64_t last_o = 0;  // observation label of h (seeds the FeedWhenCrying rollout)
  int depth = 0;
  bool want_rollout = false, aborted = false;
  long long t0 = clock64(), t_ins = 0;
  // Generative step of every action, one per lane (src/solver.jl:126 for whichever action UCB
  // will pick): it depends on the state and on this level's draws only, not on the node's
  // statistics, so it runs while those are in flight and is off the critical path.
  State sp_l;
  // observation label: 32 bits are enough unless observations are continuous (LightDark: the bits of a double)
  using ObsT = typename std::conditional<M::HASHED, uint64_t, uint32_t>::type;
  ObsT o_l;
  double r_l;
  // One Philox block feeds several levels (batched mode keeps its own stream map: block = level / LV):
  // RockSample needs one word per level (the sensor draw), the other models two.
  constexpr int TW = (PID == POMCP_ROCKSAMPLE) ? 1 : 2, LV = 4 / TW;
  Philox4 ptb;
  int a_new = 0;  // action widening: the action this level would add
  // observation widening: one block per level (word 2 draws the event).  DPO: the kernel instance carries the
  // widening code at all - always for continuous observations (as before), for discrete ones only when it is asked
  // for (a run-time flag alone cost the plain RockSample search 14 %: registers and instructions of a path it never takes)
  const bool dpw = DPO && cfg.dpw != 0;
  // The tree draws of a whole simulation in one lane-parallel Philox call: lane j holds block j (levels
  // j * LV .. j * LV + LV - 1); a level fetches its words with a shuffle.  Same counters, same words as
  // computing block d / LV at level d (which levels past the 32nd block still do).
  Philox4 pall{};
  if (!dpw) pall = philox4x32_10((uint32_t)lane, TAG_TREE | tag_rank, q, tv.epoch, tv.key0, tv.key1);
  auto spec_step = [&](int d) {
    uint32_t w0, w1 = 0;
    if (dpw) {
      ptb = philox4x32_10((uint32_t)d, TAG_TREE | tag_rank, q, tv.epoch, tv.key0, tv.key1);
      w0 = ptb.w[0];
      w1 = ptb.w[1];
    } else {
      const int blk = d / LV, k = d % LV;
      if (blk < 32) {
        if (TW == 1) {
          const uint32_t lo = (k & 1) ? pall.w[1] : pall.w[0], hi = (k & 1) ? pall.w[3] : pall.w[2];
          w0 = __shfl_sync(0xffffffffu, (k & 2) ? hi : lo, blk);
        } else {
          w0 = __shfl_sync(0xffffffffu, k ? pall.w[2] : pall.w[0], blk);
          w1 = __shfl_sync(0xffffffffu, k ? pall.w[3] : pall.w[1], blk);
        }
      } else {
        const Philox4 pb2 = philox4x32_10((uint32_t)blk, TAG_TREE | tag_rank, q, tv.epoch, tv.key0, tv.key1);
        if (TW == 1) {
          w0 = k == 0 ? pb2.w[0] : k == 1 ? pb2.w[1] : k == 2 ? pb2.w[2] : pb2.w[3];
        } else {
          w0 = k ? pb2.w[2] : pb2.w[0];
          w1 = k ? pb2.w[3] : pb2.w[1];
        }
      }
    }
    if (DPA)  // next_action(RandomActionGenerator): rand(rng, actions(pomdp)), its own block per level
      a_new = (int)__umulhi(philox4x32_10((uint32_t)d | 0x80000000u, TAG_TREE | tag_rank, q, tv.epoch, tv.key0, tv.key1).w[0],
                            (uint32_t)nA);
    const int a_l = lane < nA ? lane : nA - 1;
    uint64_t o64;
    if constexpr (PID == POMCP_ROCKSAMPLE) rs_tree_step(m, tb, eta_thr, s, a_l, w0, sp_l, o64, r_l);
    else M::template step<true, double>(m, s, a_l, u01(w0), M::T_DRAW ? u01(w1) : u01(w0), u01(w1), sp_l, o64, r_l);
    o_l = (ObsT)o64;
  };
  spec_step(0);
  for (;;) {
    if (depth >= cfg.levels || M::terminal(m, s)) break;  // src/solver.jl:82-86
#ifdef POMCP_PROBE
    const long long tp0 = clock64();
    asm volatile("" ::"r"(st.n), "d"(st.w), "r"(st.mm), "r"(st.kid[0]), "r"(st.kid[ActStats<PID>::K - 1]));
    const long long tpw = clock64();
    lc.cyc[4] += (unsigned long long)(tpw - tp0);  // waiting for the statistics
#endif
    uint32_t amask = 0xffffffffu;
    if (DPA) {
      // action progressive widening, src/solver.jl:172-186: total_N = sum of the children's N (in-flight visits
      // included), a uniformly drawn action is added while #children <= k * total_N^alpha, and a node left with at
      // most one child returns its leaf estimate.  The mask read with the statistics may be a few additions
      // behind; the bit is added with a fire-and-forget atomic.
      amask = st.amask;
      const uint32_t total = __reduce_add_sync(0xffffffffu, (lane < nA && ((amask >> lane) & 1u)) ? st.n : 0u);
      hN = total;
      if (dpw_widens_actions(cfg, total, __popc(amask))) {
        const uint32_t bit = 1u << a_new;
        if (!(amask & bit)) {
          amask |= bit;
          if (lane == 0) atomicOr(&P.node_amask[h], bit);
        }
        if (__popc(amask) <= 1) {
          want_rollout = depth > 0;
          break;
        }
      }
    } else if (!(fl & 1u)) {  // src/solver.jl:87-107: the first visitor that is not cut off expands the node
      int i_expand = 0;
      if (lane == 0) i_expand = !(atomicOr(&P.node_flags[h], 1u) & 1u);
      i_expand = __shfl_sync(0xffffffffu, i_expand, 0);
      if (i_expand) {
        want_rollout = depth > 0;
        break;
      }
    }
    // UCB1, src/solver.jl:109-124 (lanes = actions)
    // Straight-line on every lane (selects, no branches): rcp(0) = +inf makes an untried action's bonus +inf
    // by itself; the two special cases of the reference are patched in afterwards.
    uint32_t key;
    {
      const uint32_t hn = hN < 1u ? 1u : hN;
      const float v = (st.mm > 0 && v_finite) ? (float)st.w * sfu_rcp((float)st.mm) : initV_f;
      float crit = v + c_f * sfu_sqrt(sfu_lg2((float)hn) * 0.69314718f * sfu_rcp((float)st.n));
      const bool n0 = st.n == 0;
      crit = (n0 && v == -INFINITY) ? INFINITY : crit;
      crit = (n0 && hn == 1u) ? v : crit;  // src/solver.jl:112 / :210
      const bool valid = lane < nA && ((amask >> lane) & 1u);
      key = (valid && crit == crit) ? f2ord(crit) : 0u;
    }
    const uint32_t mx = __reduce_max_sync(0xffffffffu, key);
    const unsigned tied = __ballot_sync(0xffffffffu, key == mx && lane < nA);
    const int best_a = mx ? 31 - __clz((int)tied) : nA - 1;  // ">=": the last maximal index wins
    const uint32_t slot = h * (uint32_t)nA + (uint32_t)best_a;
#ifdef POMCP_PROBE
    const long long tp1 = clock64();
    lc.cyc[5] += (unsigned long long)(tp1 - tpw);  // UCB + arg-max
#endif

    // outcome of the chosen action; child ids came with the statistics (ids never change once published)
    uint32_t mine = 0;
    if (!M::HASHED) {
      mine = st.kid[0];
#pragma unroll
      for (int k = 1; k < ActStats<PID>::K; ++k) mine = (o_l == (ObsT)k) ? st.kid[k] : mine;
    }
    uint32_t raw = __shfl_sync(0xffffffffu, mine, best_a);
    State sp = shfl_state(sp_l, best_a);
    const uint64_t o = (uint64_t)__shfl_sync(0xffffffffu, o_l, best_a);
    const double r = __shfl_sync(0xffffffffu, r_l, best_a);
    // observation progressive widening (src/solver.jl:223-262): re-use an earlier generated (child, state)
    bool reuse = false;
    if (dpw) {
      const uint32_t bn = __shfl_sync(0xffffffffu, st.n, best_a);
      const uint32_t nc = __shfl_sync(0xffffffffu, st.nc, best_a), ng = __shfl_sync(0xffffffffu, st.ng, best_a);
      if (ng > 0 && !dpw_generates(cfg, bn, nc)) {
        uint32_t node = 0;
        State spe = sp;
        int ok = 0;
        if (lane == 0) {
          const uint32_t e = __umulhi(ptb.w[2], ng);  // floor(u * #events)
          ok = ev_lookup<State>(P, slot, e, node, spe) ? 1 : 0;  // a miss = the event is still being published
        }
        ok = __shfl_sync(0xffffffffu, ok, 0);
        if (ok) {
          reuse = true;
          raw = __shfl_sync(0xffffffffu, node, 0);
          sp = shfl_state(spe, 0);  // r = reward(s, a, sp) does not depend on sp in these models
        }
      }
    }
    const bool cut_next = depth + 1 >= cfg.levels || M::terminal(m, sp);

    // src/solver.jl:132-142
    int created = 0;
    if (raw == 0) {
      const long long ti = clock64();
      if (lane == 0) {
        bool cr;
        if constexpr (M::HASHED) raw = find_or_insert_child<PID, false>(P, m, slot, o, cr, false);
        else raw = insert_child_spare(P, m.nO, slot, o, cr, !cut_next && !DPA, spare,
                                      (M::TWO_STATE && cfg.exact2) ? bayes2<PID>(m, __ldcg(&P.node_bel[h]), best_a, o) : 0.0);
        created = cr;
      }
      raw = __shfl_sync(0xffffffffu, raw, 0);
      created = __shfl_sync(0xffffffffu, created, 0);
      t_ins += clock64() - ti;
      if (raw == 0) { aborted = true; break; }
    }
    const uint32_t hao = raw & CHILD_ID_MASK;
    const bool i_expand_child = created && (raw & CHILD_EXPANDED);  // direct table: the creator expands
    const bool try_claim = created && !i_expand_child && !cut_next && !DPA;  // hashed table: claim now
    const bool flag_known = (raw & CHILD_EXPANDED) != 0;
    // Arrival counts, then everything the next level waits for — the child's visit count, expansion flag
    // and statistics — in one round trip, requested before any bookkeeping.  The count and the flag are read
    // by every lane (one broadcast transaction, no shuffle, no divergence); whether the count includes this
    // worker's own arrival does not matter to log(h.N).
    const bool l0 = lane == 0;
    red_add_u32_if(l0, &P.act_N[slot]);
    red_add_u32_if(l0 && !reuse, &P.node_N[hao]);  // hao.N counts generated arrivals only under widening
    uint32_t seen = 0, nfl = flag_known ? 1u : 0u;
    if (!i_expand_child) {
      if (!try_claim) st.load(P, hao, nA, lane, DPO && cfg.dpw != 0, DPA);
      seen = __ldcg(&P.node_N[hao]);
      if (try_claim) {  // hashed table: bit1 = the claim was attempted here
        if (lane == 0) nfl = atomicOr(&P.node_flags[hao], 1u) | 2u;
        nfl = __shfl_sync(0xffffffffu, nfl, 0);
      } else if (!flag_known) {
        nfl = __ldcg(&P.node_flags[hao]);
      }
    }
#ifdef POMCP_PROBE
    const long long tp2 = clock64();
    lc.cyc[6] += (unsigned long long)(tp2 - tp1);  // pick + request the child
#endif
    // speculative step of the next level, then the bookkeeping of this one: both run under the loads
    s = sp;
    if (!i_expand_child && !cut_next) spec_step(depth + 1);
    if (dpw && !reuse) {
      if (lane == 0) {
        if (created) atomicAdd(&P.act_nc[slot], 1u);
        ev_insert<State>(P, slot, atomicAdd(&P.act_ng[slot], 1u), hao, sp);
      }
      __syncwarp();
    }
    red_or_u32_if(l0 && i_expand_child, &P.node_flags[hao], 1u);  // nobody waits for it
    const int pi = depth * pv.stride;
    st_u32_if(l0, &pv.slot[pi], slot);
    st_f64_if(l0, &pv.r[pi], r);
    if (cfg.belief_unweighted) {
      if (lane == 0) {
        pv.node[pi] = hao;
        pv.state[pi] = sp;
      }
      // re-converge: a warp left split here runs everything up to the next shuffle twice (measured once: the
      // whole rollout loop at 15.5 active lanes)
      __syncwarp();
    }
#ifdef POMCP_PROBE
    lc.cyc[7] += (unsigned long long)(clock64() - tp2);  // speculative step + bookkeeping
#endif
    h = hao;
    last_o = o;
    ++depth;
    if (i_expand_child) {  // expand + evaluate the leaf (src/solver.jl:87-103)
      want_rollout = true;
      break;
    }
    hN = seen;
    // The flag races with other workers' claims and steers control flow at the top of the next level: every lane
    // must see the same value.  `flag_known` is warp-uniform (it comes out of a shuffle) and the claimed flag was
    // broadcast above; only the per-lane load needs lane 0's copy.
    fl = (flag_known || try_claim || i_expand_child) ? nfl : __shfl_sync(0xffffffffu, nfl, 0);
    if (M::HASHED && (fl & 2u)) {  // (the direct table never claims here: its creator expands)
      if (!(fl & 1u)) {  // claimed
        want_rollout = true;
        break;
      }
      st.load(P, hao, nA, lane, DPO && cfg.dpw != 0, DPA);  // somebody else got there first (rare): descend through it
      fl = 1u;
    }
  }
  if (lane == 0) {
    if (aborted) atomicMax(&tv.ctl->next_query, 1ull << 62);  // pool exhausted: stop every worker of this tree
    next_q = atom_add_u64_async(&tv.ctl->next_query, 1ull, P.max_nodes);
  }
  __syncwarp();  // lanes re-converge before the rollouts (see above)
  if (aborted) return;
  const long long t1 = clock64();
  lc.cyc[0] += (unsigned long long)(t1 - t0 - t_ins);
  lc.cyc[1] += (unsigned long long)t_ins;

  // leaf evaluation, src/solver.jl:97-103
  double leaf = 0.0;
  if (want_rollout) {
    long long steps_to_eps = (long long)ceil(cfg.log_ratio - (double)depth);
    long long steps = min(cfg.max_depth - (long long)depth, steps_to_eps);
    if (cfg.dpw_solver) steps = depth;  // estimate_value(..., s, h, depth), src/solver.jl:182,199
    double est;
    if (cfg.estimator == POMCP_EST_CONSTANT) {
      est = cfg.est_value;  // src/rollout.jl:10
      lc.rollouts++;
    } else if (cfg.estimator == POMCP_EST_STATE_VALUE) {  // value(policy, s), src/rollout.jl:67-69
      const long long si = state_index(s);
      est = (M::terminal(m, s) || si < 0 || si >= cfg.n_state_values) ? 0.0 : cfg.state_values[si];
      lc.rollouts++;
    } else {
      const int W = cfg.team_warps;
      if (W > 1) {  // post the leaf to the helper warps
        if (lane == 0) {
          job->s = s;
          job->steps = steps;
          job->last_o = last_o;
          job->q = q;
          job->cmd = 1;
        }
        team_bar<NTEAMS>(job->bar, 32 * W);
      }
      int nsteps = 0;
      const RngKey rk{tv.key0, tv.key1, tv.tagrank, tv.epoch};
      const double ret = leaf_rollouts<PID>(m, tb, cfg, rk, s, steps, last_o, q, lane, nsteps);
      double sum = warp_tree_sum(ret);
      nsteps = __reduce_add_sync(0xffffffffu, nsteps);
      if (W > 1) {
        team_bar<NTEAMS>(job->bar + 1, 32 * W);  // the helpers' partial sums are in shared memory
        for (int w = 1; w < W; ++w) {
          sum += job->partial[w];
          nsteps += job->psteps[w];
        }
      }
      est = sum / (double)cfg.rollouts;
      lc.rollouts += cfg.rollouts;
      lc.rollout_steps += nsteps;
    }
    leaf = cfg.gpow[depth] * est;  // src/solver.jl:103 (the reference's extra gamma^depth is kept)
  }
  const long long t2 = clock64();
  lc.cyc[2] += (unsigned long long)(t2 - t1);

  // particle pushes, deepest level first like the unwinding recursion (src/solver.jl:146-148)
  if (cfg.belief_unweighted && depth > 0) {
    unsigned long long log_base = 0;
    if (lane == 0) {
      log_base = atomicAdd(tv.log_ctr, (unsigned long long)depth);
      if (log_base + depth > tv.log_cap) { P.ctr->log_overflow = 1; log_base = ~0ull; }
      else log_base += tv.log_base;
    }
    log_base = __shfl_sync(0xffffffffu, log_base, 0);
    if (log_base != ~0ull) {
      State* log_state = (State*)P.log_state;
      for (int d = depth - 1 - lane; d >= 0; d -= 32) {
        const unsigned long long li = log_base + (unsigned long long)(depth - 1 - d);
        P.log_node[li] = pv.node[d * pv.stride];
        log_state[li] = pv.state[d * pv.stride];
      }
    }
  }
  // backup, src/solver.jl:144-156: W += R, M += 1 (V = W / M) with R_d = r_d + gamma * R_{d+1}, R_depth = leaf.
  // Lane-parallel: lane l of a 32-level chunk holds level lo + l; the discounted suffix sums come out of a
  // five-step warp scan (v_d += gamma^k * v_{d+k}, k = 1, 2, 4, 8, 16) and every lane publishes its own level.
  {
    const double g1 = m.gamma, g2 = g1 * g1, g4 = g2 * g2, g8 = g4 * g4, g16 = g8 * g8;
    double Rnext = leaf;
    for (int hi = depth; hi > 0; hi -= 32) {  // deepest chunk first
      const int lo = hi > 32 ? hi - 32 : 0, n = hi - lo;
      const bool act = lane < n;
      const int pi = (lo + lane) * pv.stride;
      const double gtail = act ? cfg.gpow[n - lane] : 0.0;  // gamma^(levels between this one and R_hi)
      double v = act ? pv.r[pi] : 0.0;
      const uint32_t slot = act ? pv.slot[pi] : 0u;
      double t;
      t = __shfl_down_sync(0xffffffffu, v, 1);  if (lane + 1 < 32) v = v + g1 * t;
      t = __shfl_down_sync(0xffffffffu, v, 2);  if (lane + 2 < 32) v = v + g2 * t;
      t = __shfl_down_sync(0xffffffffu, v, 4);  if (lane + 4 < 32) v = v + g4 * t;
      t = __shfl_down_sync(0xffffffffu, v, 8);  if (lane + 8 < 32) v = v + g8 * t;
      t = __shfl_down_sync(0xffffffffu, v, 16); if (lane + 16 < 32) v = v + g16 * t;
      const double R = v + gtail * Rnext;
      if (act) {
        atomicAdd(&P.act_V[slot], R);
        atomicAdd(&P.act_M[slot], 1u);
      }
      Rnext = __shfl_sync(0xffffffffu, R, 0);
    }
    if (lane == 0) atomicAdd(&P.node_N[tv.root], 1u);  // b.N += 1 after simulate (src/solver.jl:51)
  }
  __syncwarp();
  lc.cyc[3] += (unsigned long long)(clock64() - t2);
  lc.sims++;
  lc.tree_steps += depth;
  if ((unsigned long long)depth > lc.max_depth) lc.max_depth = depth;
}

__device__ __forceinline__ void flush_counters(const Pool& P, const LocalCtr& lc, bool exact) {
  DevCtr* c = P.ctr;
  if (exact) {
    c->sims += lc.sims; c->skips += lc.skips; c->rollouts += lc.rollouts; c->rollout_steps += lc.rollout_steps;
    c->tree_steps += lc.tree_steps;
    if (lc.max_depth > c->max_depth) c->max_depth = lc.max_depth;
    if (lc.sims) c->any_nonterminal = 1;
  } else {
    if (lc.sims) { atomicAdd(&c->sims, lc.sims); atomicAdd(&c->tree_steps, lc.tree_steps); c->any_nonterminal = 1; }
    if (lc.skips) atomicAdd(&c->skips, lc.skips);
    if (lc.rollouts) { atomicAdd(&c->rollouts, lc.rollouts); atomicAdd(&c->rollout_steps, lc.rollout_steps); }
    atomicMax(&c->max_depth, lc.max_depth);
    for (int i = 0; i < 4; ++i)
      if (lc.cyc[i]) atomicAdd(&c->dbg[i], lc.cyc[i]);
#ifdef POMCP_PROBE  // the probe build reports its four descent segments in dbg[4..7]
    for (int i = 4; i < 8; ++i) atomicAdd(&c->dbg[i], lc.cyc[i]);
#endif
  }
}

// ---- batched mode: persistent asynchronous warp workers, one warp per CTA.  A worker repeatedly
// takes the next query id, runs one whole simulation and publishes its statistics with atomics.
// There is no barrier anywhere: the number of simulations in flight is the number of admitted
// workers, clamp(root visits / ramp_div, inflight_min, inflight_max), so the tree is never more than
// that many simulations stale (small while the tree is small).  Root visits only grow, so a worker
// waits for admission once: worker w starts when root visits >= (w + 1) * ramp_div.
// TEAM = warps per worker (1, 2 or 4), NTEAMS = 4 / TEAM workers per CTA: CTAs are always four warps, one
// per scheduler of the SM.  Launch bounds keep 1024 workers co-resident on 148 SMs (16 warps of <= 128
// registers per SM; the four-warp teams run 7 CTAs per SM at 72 registers).
// Measured oddity (scripts/cycle_probe.py, same rollout code as a __noinline__ function): the 32
// rollouts of the tree warp take 18.4 k cycles when it rolls out alone (TEAM = 1) and 12.1 k when a
// helper warp rolls out next to it (TEAM = 2), so R = 64 is not only free, it is faster per simulation.
template <int PID, int TEAM, int NTEAMS, bool DPA, bool DPO = ModelT<PID>::HASHED>
__global__ void __launch_bounds__(32 * TEAM * NTEAMS, TEAM == 4 ? 7 : 4) k_search_workers(
    const __grid_constant__ Pool P, const __grid_constant__ DevModel m, const __grid_constant__ SearchCfg cfg,
    const __grid_constant__ PathBuf pb, long long queries) {
  using State = typename ModelT<PID>::State;
  extern __shared__ __align__(16) unsigned char dyn_smem[];
  __shared__ RsTab rs_tab;
  __shared__ LeafJob<State> jobs[NTEAMS];
  __shared__ TreeView tvs[NTEAMS];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int team = warp / TEAM, role = warp % TEAM;
  constexpr int W = TEAM;
  // tables shared by the CTA
  if (PID == POMCP_ROCKSAMPLE) {
    uint32_t* e = (uint32_t*)(dyn_smem + cfg.eta_off);
    for (int i = threadIdx.x; i < cfg.eta_n; i += blockDim.x) e[i] = cfg.eta_thr[i];
    rs_tab_init(&rs_tab, m);  // ends with __syncthreads()
  }
  const uint32_t* eta_thr = (const uint32_t*)(dyn_smem + cfg.eta_off);
  LeafJob<State>* job = &jobs[team];
  const TreeView& tv = tvs[team];
  const int bar = 1 + 2 * team;
  if (role > 0) {
    // helper warp of the rollout team: 32 rollouts of every leaf the tree warp posts
    for (;;) {
      team_bar<NTEAMS>(bar, 32 * W);
      if (job->cmd == 0) break;
      int nsteps = 0;
      const RngKey rk{tv.key0, tv.key1, tv.tagrank, tv.epoch};  // written by the tree warp before its first barrier
      const double ret = leaf_rollouts<PID>(m, &rs_tab, cfg, rk, job->s, job->steps, job->last_o, job->q, lane, nsteps, 32 * role);
      const double sum = warp_tree_sum(ret);
      nsteps = __reduce_add_sync(0xffffffffu, nsteps);
      if (lane == 0) {
        job->partial[role] = sum;
        job->psteps[role] = nsteps;
      }
      team_bar<NTEAMS>(bar + 1, 32 * W);
    }
    return;
  }
  DevCtr* c = P.ctr;
  // worker ids are handed out in start order, so the admitted (lowest) ids are always resident
  unsigned wid = 0;
  if (lane == 0) wid = (unsigned)atomicAdd(&c->worker_ticket, 1ull);
  wid = __shfl_sync(0xffffffffu, wid, 0);
  // the worker's tree: ids are dealt round-robin, wt = its index among the workers of that tree
  const unsigned ntr = cfg.n_trees > 1 ? (unsigned)cfg.n_trees : 1u;
  const unsigned ti = wid % ntr, wt = wid / ntr;
  TreeCtl* const tc = cfg.trees + ti;
  if (lane == 0) {
    job->bar = bar;
    TreeView& w = tvs[team];
    w.bel = tc->bel; w.n_bel = tc->n_bel; w.ctl = tc;
    w.root = tc->root; w.root_N0 = tc->root_N0;
    w.key0 = tc->key0; w.key1 = tc->key1; w.tagrank = tc->rank << 8; w.epoch = tc->epoch;
    w.log_ctr = tc->log_ctr; w.log_base = tc->log_base; w.log_cap = tc->log_cap;
  }
  __syncwarp();
  PathView<State> pv;
  if (cfg.path_smem) {
    const int L = cfg.levels > 0 ? cfg.levels : 1;
    unsigned char* base = dyn_smem + (size_t)team * cfg.path_bytes;
    pv.r = (double*)base;
    pv.state = (State*)(base + (size_t)L * 8);
    pv.slot = (uint32_t*)(base + (size_t)L * (8 + sizeof(State)));
    pv.node = pv.slot + L;
    pv.stride = 1;
  } else {
    pv.r = pb.path_r + wid;
    pv.state = (State*)pb.path_state + wid;
    pv.slot = pb.path_slot + wid;
    pv.node = pb.path_node + wid;
    pv.stride = pb.stride;
  }
  LocalCtr lc{};
  unsigned long long next_q = ~0ull;
  uint32_t spare = 0;  // lane 0: node id allocated ahead of the next insertion (direct child table)
  const long long t_start = clock64();
  if (lane == 0 && (long long)wt < (long long)cfg.inflight_max) {
    // admission: ramp the number of simulations in flight with the size of the tree (cfg.inflight_* are per tree)
    const unsigned long long need =
        (long long)wt < (long long)cfg.inflight_min ? 0ull : ((unsigned long long)wt + 1ull) * (unsigned long long)cfg.ramp_div;
    bool go = true;
    const uint32_t root = tc->root;
    unsigned long long t_wait0 = 0;
    for (;;) {
      const unsigned long long seen = (unsigned long long)__ldcg(&P.node_N[root]);
      if (seen >= need) break;
      if (__ldcg((const unsigned long long*)&tc->next_query) >= (unsigned long long)queries) { go = false; break; }
      unsigned long long now;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(now));
      if (t_wait0 == 0) t_wait0 = now;
      if (now - t_wait0 > 2000000000ull) {  // 2 s without being admitted: give up instead of hanging the device
        c->pool_overflow = 2;
        go = false;
        break;
      }
      // Sleep in proportion to the visits still missing: the tree cannot grow faster than about one simulation per
      // 8 ns, and a polling lane takes the issue slots of a whole warp (a fixed 2 us sleep - which the hardware cuts
      // short - made the admission loop 5 % of all warp instructions of a 1e6-query search).
      const unsigned long long gap = need - seen;
      __nanosleep((unsigned)(gap * 8ull < 1000ull ? 1000ull : (gap * 8ull > 100000ull ? 100000ull : gap * 8ull)));
    }
    if (go) next_q = atomicAdd(&tc->next_query, 1ull);
  }
  const long long t_admitted = clock64();
  for (;;) {
    const unsigned long long q = __shfl_sync(0xffffffffu, next_q, 0);
    if (q >= (unsigned long long)queries) break;
    simulate_async<PID, DPA, NTEAMS, DPO>(P, m, &rs_tab, eta_thr, cfg, tv, pv, job, (uint32_t)q, lc, next_q, spare);
  }
  if (W > 1) {  // release the helpers
    if (lane == 0) job->cmd = 0;
    team_bar<NTEAMS>(bar, 32 * W);
  }
  if (lane == 0) {
    flush_counters(P, lc, false);
    if (lc.sims) tc->any_nonterminal = 1;
#ifndef POMCP_PROBE
    atomicAdd(&c->dbg[4], (unsigned long long)(t_admitted - t_start));  // waiting for admission
    atomicAdd(&c->dbg[5], (unsigned long long)(clock64() - t_admitted));  // working
#endif
  }
}

// ---- exact mode: one warp runs the queries strictly one after another (bit-parity with the
// sequential reference; also the replay hook).
template <int PID, bool REPLAY>
__global__ void __launch_bounds__(32) k_search_exact(const __grid_constant__ Pool P,
                                                     const __grid_constant__ DevModel m,
                                                     const __grid_constant__ SearchCfg cfg,
                                                     const __grid_constant__ PathBuf pb, long long queries,
                                                     SearchResult* res) {
  __shared__ RsTab rs_tab;
  if (PID == POMCP_ROCKSAMPLE) rs_tab_init(&rs_tab, m);
  ReplayCursor rc{0, 0};
  LocalCtr lc{};
  uint32_t spare = 0;
  const int lane = threadIdx.x & 31;
  for (long long q = 0; q < queries; ++q) {
    uint32_t root_ticket = __ldcg(&P.node_N[cfg.root]);
    simulate_one<PID, true, REPLAY>(P, m, &rs_tab, cfg, pb, 0, (uint32_t)q, root_ticket, rc, spare, lc);
    __threadfence_block();
    __syncwarp();
  }
  if (lane == 0) {
    flush_counters(P, lc, true);
    res->ucur = rc.ucur;
    res->rcur = rc.rcur;
    if (REPLAY && (rc.ucur > cfg.n_uni || rc.rcur > cfg.n_rets)) P.ctr->replay_overflow = 1;
  }
}

// ---- K5: root statistics + arg-max V with ">=" (src/solver.jl:60-70); also packs the
// [N_a, N_a*V_a, exists_a] vector for the root-parallel merge (K10).  The reference iterates
// values(b.children): under action widening only the ActNodes that were actually added take part
// (a slot that was never added still holds the pool's prefill and must not win a ">=" against
// negative values).
__global__ void k_root_result(const __grid_constant__ Pool P, int nA, const TreeCtl* trees, int exact, double init_V,
                              SearchResult* res_all) {
  if (threadIdx.x != 0) return;
  const int t = blockIdx.x;  // one block per tree
  SearchResult* res = res_all + t;
  const uint32_t root = trees[t].root;
  DevCtr* c = P.ctr;
  res->root_N = P.node_N[root];
  const uint32_t amask = P.node_amask ? P.node_amask[root] : 0xffffffffu;
  int best = -1;
  double best_V = 0.0;
  bool nan_seen = false;
  for (int a = 0; a < nA; ++a) {
    uint32_t slot = root * (uint32_t)nA + a;
    double v = P.act_V[slot];
    long long n = P.act_N[slot];
    if (!exact) {
      uint32_t mm = P.act_M[slot];
      v = (mm > 0 && fabs(init_V) != INFINITY) ? v / (double)mm : init_V;
    }
    const bool exists = (amask >> a) & 1u;
    res->N[a] = exists ? n : 0;
    res->V[a] = v;
    res->merge[a] = exists ? (double)n : 0.0;
    res->merge[nA + a] = exists ? (double)n * v : 0.0;
    res->merge[2 * nA + a] = exists ? 1.0 : 0.0;
    if (!exists) continue;
    if (best < 0) nan_seen = v != v;  // best_V starts from the first child (src/solver.jl:61-62)
    if (best < 0 || v >= best_V) { best_V = v; best = a; }
  }
  res->action = best < 0 ? 0 : best;
  int st = POMCP_OK;
  if (nan_seen) st = POMCP_ERR_NAN_VALUE;                          // @assert !isnan(best_V), src/solver.jl:62
  if (!(exact ? c->any_nonterminal : trees[t].any_nonterminal)) st = POMCP_ERR_ALL_SAMPLES_TERMINAL;  // src/solver.jl:56-58
  if (c->pool_overflow || c->log_overflow) st = POMCP_ERR_POOL_EXHAUSTED;
  if (c->pool_overflow == 2) st = POMCP_ERR_CUDA;  // worker watchdog fired
  if (c->replay_overflow) st = POMCP_ERR_INVALID_ARG;
  res->status = st;
}

// Several trees on one device (root parallelism inside a GPU): the same merge as K10, over the trees of this
// handle.  res_all[0..T-1] are the trees' own results, res_all[T] receives the merged one; a tree whose samples
// were all terminal contributes nothing (it has no children), any other failure is the merged status.
__global__ void k_trees_merge(int T, int nA, double init_V, SearchResult* res_all) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  SearchResult* out = res_all + T;
  int st = POMCP_ERR_ALL_SAMPLES_TERMINAL;
  long long root_N = 0;
  for (int i = 0; i < 3 * nA; ++i) out->merge[i] = 0.0;
  for (int t = 0; t < T; ++t) {
    const SearchResult* r = res_all + t;
    root_N += r->root_N;
    if (r->status == POMCP_ERR_ALL_SAMPLES_TERMINAL) continue;
    if (r->status != POMCP_OK) { st = r->status; break; }
    if (st == POMCP_ERR_ALL_SAMPLES_TERMINAL) st = POMCP_OK;
    for (int i = 0; i < 3 * nA; ++i) out->merge[i] += r->merge[i];  // fixed order: trees 0, 1, 2, ...
  }
  out->status = st;
  out->root_N = root_N;
  out->ucur = out->rcur = 0;
  int best = -1;
  double best_V = 0.0;
  for (int a = 0; a < nA; ++a) {
    const double n = out->merge[a], nv = out->merge[nA + a];
    const double v = n > 0.0 ? nv / n : init_V;
    out->N[a] = (long long)n;
    out->V[a] = v;
    if (!(out->merge[2 * nA + a] > 0.0)) continue;
    if (best < 0 || v >= best_V) { best_V = v; best = a; }
  }
  out->action = best < 0 ? 0 : best;
}

// after the all-reduce: V_a = sum(N*V)/sum(N) over the trees that hold ActNode a, then the reference's arg-max rule
__global__ void k_root_merge_finish(int nA, double init_V, SearchResult* res) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  int best = -1;
  double best_V = 0.0;
  for (int a = 0; a < nA; ++a) {
    double n = res->merge[a], nv = res->merge[nA + a];
    const bool exists = res->merge[2 * nA + a] > 0.0;
    double v = n > 0.0 ? nv / n : init_V;
    res->N[a] = (long long)n;
    res->V[a] = v;
    if (!exists) continue;
    if (best < 0 || v >= best_V) { best_V = v; best = a; }
  }
  res->action = best < 0 ? 0 : best;
}

// ---- pool (re)initialisation: ActNode(a, init_N, init_V, {}) for every slot (src/solver.jl:89-93)
__global__ void k_pool_fill(const __grid_constant__ Pool P, int nA, int nO, unsigned long long n_nodes,
                            uint32_t init_N, double init_V, int exact) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  double v0 = exact ? init_V : (init_N > 0 && fabs(init_V) != INFINITY ? init_V * (double)init_N : 0.0);
  for (unsigned long long k = i; k < n_nodes * nA; k += stride) {
    P.act_N[k] = init_N;
    P.act_M[k] = init_N;
    P.act_V[k] = v0;
    if (P.act_nc) { P.act_nc[k] = 0; P.act_ng[k] = 0; }
  }
  for (unsigned long long k = i; k < n_nodes; k += stride) {
    P.node_N[k] = 0;
    P.node_flags[k] = 0;
    P.node_pslot[k] = 0;
    if (P.node_obs) P.node_obs[k] = 0;
    if (P.node_amask) P.node_amask[k] = 0;
  }
  if (P.child)
    for (unsigned long long k = i; k < n_nodes * nA * nO; k += stride) P.child[k] = 0;
}
__global__ void k_set_node_bel(double* node_bel, uint32_t node, const double* src, double value) {
  if (threadIdx.x == 0 && blockIdx.x == 0) node_bel[node] = src ? *src : value;
}
// update(node_belief_updater, b_old.B, a, o) for a child the tree does not hold (src/updater.jl:33)
template <int PID>
__global__ void k_bayes2(const __grid_constant__ DevModel m, const double* node_bel, uint32_t node, int a,
                         unsigned long long obs, double* out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *out = bayes2<PID>(m, node_bel[node], a, obs);
}

// ---- inspection / update helpers (single thread)
struct NodeQuery {
  int depth;
  int actions[64];
  unsigned long long obs[64];
};
struct NodeStats {
  long long node;  // -1: absent
  double bel;      // exact two-state belief of the node (EXACT2)
  long long N, n_particles;
  long long act_N[POMCP_MAX_ACTIONS];
  double act_V[POMCP_MAX_ACTIONS];
};

template <int PID>
__global__ void k_node_stats(const __grid_constant__ Pool P, const __grid_constant__ DevModel m, uint32_t root,
                             const __grid_constant__ NodeQuery qy, int exact, double init_V, long long n_bel,
                             int unweighted, NodeStats* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  uint32_t h = root;
  for (int d = 0; d < qy.depth; ++d) {
    uint32_t c = find_child<PID>(P, m, h * (uint32_t)m.nA + (uint32_t)qy.actions[d], qy.obs[d]);
    if (c == 0) { out->node = -1; return; }
    h = c;
  }
  out->node = h;
  out->bel = P.node_bel ? P.node_bel[h] : 0.0;
  out->N = P.node_N[h];
  out->n_particles = (h == root) ? n_bel : (unweighted ? (long long)P.node_N[h] : 0);
  for (int a = 0; a < m.nA; ++a) {
    uint32_t slot = h * (uint32_t)m.nA + a;
    double v = P.act_V[slot];
    if (!exact) {
      uint32_t mm = P.act_M[slot];
      v = (mm > 0 && fabs(init_V) != INFINITY) ? v / (double)mm : init_V;
    }
    out->act_N[a] = P.act_N[slot];
    out->act_V[a] = v;
  }
}

}  // namespace pomcp

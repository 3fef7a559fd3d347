// Device-resident batched evaluation episodes: E independent runs of the outer
// simulation loop - POMDPToolbox.RolloutSimulator.simulate as test_solver (test/runtests.jl:19) and the est_reward
// loop of notebooks/Sanity_Checks.ipynb:47-77 drive it -
//
//     while disc > eps && !isterminal(s) && step <= max_steps
//         a = action(policy, b);  (sp, o, r) = generate_sor(pomdp, s, a, rng)
//         total += disc * r;  s = sp;  b = update(updater, b, a, o);  disc *= gamma;  step += 1
//
// with one search tree per episode, all E trees searched side by side in ONE launch of the worker kernel per
// step (tree.cuh: n_trees), the environment step and the belief update on the device, and no host round trip
// inside an episode: the host only enqueues max_steps rounds of kernels and reads the rewards at the end.
// Everything that changes from step to step lives in the per-tree TreeCtl table and the EpisodeCtl table below.
#pragma once
#include <cstdint>
#include "pf.cuh"
#include "tree.cuh"

namespace pomcp {

struct EpisodeCtl {
  double total, disc;            // discounted reward so far, gamma^steps
  unsigned long long state[2];   // environment state (u32 or LDState)
  unsigned long long last_o;     // observation of the last step (bits)
  int last_a;
  int steps, status, active;     // status: POMCP_OK, or what ended the episode early (depletion, pool exhausted, ...)
  int update_pending;            // SIR / unweighted: the belief update of this step still has to run
  int _pad;
};

struct EpisodeCfg {
  int n, max_steps, initial_state, updater;  // updater = pomcp_belief_mode of the handle
  double sim_eps, initial_p1;
  uint32_t env_k0, env_k1;       // episode e draws its environment from Philox(env key + e)
  uint32_t key0, key1, rank0;    // planner streams: key of the solver, stream id rank0 + e
  long long sir_n;               // particles per episode (SIR)
  unsigned long long log_seg;    // particle-log entries per episode (unweighted)
};

template <class State>
__device__ __forceinline__ State ep_load_state(const EpisodeCtl& e) {
  State s;
  memcpy(&s, e.state, sizeof(State));
  return s;
}
template <class State>
__device__ __forceinline__ void ep_store_state(EpisodeCtl& e, const State& s) {
  e.state[0] = e.state[1] = 0ull;
  memcpy(e.state, &s, sizeof(State));
}
__device__ __forceinline__ uint32_t ep_fixed_state(uint32_t, int v) { return (uint32_t)v; }
__device__ __forceinline__ LDState ep_fixed_state(const LDState& s, int) { return s; }

// s0 ~ initial_state_distribution (or the fixed initial state of the simulator), b0 = the initial belief,
// tree e rooted at node e of a fresh pool
template <int PID>
__global__ void k_ep_init(const __grid_constant__ Pool P, const __grid_constant__ DevModel m,
                          const __grid_constant__ EpisodeCfg cfg, TreeCtl* trees, EpisodeCtl* eps,
                          unsigned long long* log_ctrs, void* bel_buf) {
  using M = ModelT<PID>;
  using State = typename M::State;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cfg.n) return;
  EpisodeCtl ec{};
  const Philox4 p = philox4x32_10(1u, TAG_ENV, 0u, 0u, cfg.env_k0 + (uint32_t)e, cfg.env_k1);
  State s = M::initial(m, u01(p.w[0]), u01(p.w[1]));
  if (cfg.initial_state >= 0 && M::TWO_STATE) s = ep_fixed_state(s, cfg.initial_state);
  ep_store_state<State>(ec, s);
  ec.total = 0.0;
  ec.disc = 1.0;
  ec.active = (cfg.max_steps > 0 && !M::terminal(m, s)) ? 1 : 0;
  ec.status = POMCP_OK;
  eps[e] = ec;
  TreeCtl tc{};
  tc.root = (uint32_t)e;
  tc.root_N0 = 0;
  tc.key0 = cfg.key0; tc.key1 = cfg.key1;
  tc.rank = cfg.rank0 + (uint32_t)e;
  tc.log_ctr = log_ctrs + e;
  tc.log_base = (unsigned long long)e * cfg.log_seg;
  tc.log_cap = cfg.log_seg;
  log_ctrs[e] = 0ull;
  tc.bel = nullptr;  // unweighted / exact: the initial belief is the distribution itself
  tc.n_bel = 0;
  if (cfg.updater == POMCP_BELIEF_EXTERNAL) {  // SIR: the belief is a particle set from the start
    tc.bel = (const char*)bel_buf + (size_t)e * (size_t)cfg.sir_n * sizeof(State);
    tc.n_bel = cfg.sir_n;
  }
  if (cfg.updater == POMCP_BELIEF_EXACT2 && P.node_bel) P.node_bel[e] = cfg.initial_p1;
  trees[e] = tc;
}

// SIR: n draws from the initial distribution per episode
template <int PID>
__global__ void k_ep_init_particles(const __grid_constant__ DevModel m, const __grid_constant__ EpisodeCfg cfg,
                                    typename ModelT<PID>::State* buf) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= (long long)cfg.n * cfg.sir_n) return;
  const uint32_t e = (uint32_t)(i / cfg.sir_n);
  const long long j = i % cfg.sir_n;
  const Philox4 p = philox4x32_10(0u, TAG_INIT, (uint32_t)j, (uint32_t)((unsigned long long)j >> 32), cfg.env_k0 + e,
                                  cfg.env_k1 ^ 0x5851f42du);
  buf[i] = ModelT<PID>::initial(m, u01(p.w[0]), u01(p.w[1]));
}

// before every search: the query tickets of the live episodes, the step's Philox epoch
__global__ void k_ep_prepare(TreeCtl* trees, const EpisodeCtl* eps, int n, uint32_t epoch) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n) return;
  trees[e].next_query = eps[e].active ? 0ull : (1ull << 62);  // a finished episode's workers leave at once
  trees[e].any_nonterminal = 0;
  trees[e].epoch = epoch;
}

// after every search: act, step the environment, account the reward, and re-root (exact / unweighted) or hand the
// particle update to the kernels that follow (SIR, unweighted gather)
template <int PID>
__global__ void k_ep_step(const __grid_constant__ Pool P, const __grid_constant__ DevModel m,
                          const __grid_constant__ EpisodeCfg cfg, TreeCtl* trees, EpisodeCtl* eps,
                          const SearchResult* res, uint32_t step) {
  using M = ModelT<PID>;
  using State = typename M::State;
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= cfg.n) return;
  EpisodeCtl& ec = eps[e];
  ec.update_pending = 0;
  if (!ec.active) return;
  const SearchResult& r = res[e];
  if (r.status != POMCP_OK) {  // AllSamplesTerminal / pool exhausted: the episode ends here with that status
    ec.status = r.status;
    ec.active = 0;
    return;
  }
  const int a = r.action;
  State s = ep_load_state<State>(ec), sp;
  const Philox4 p = philox4x32_10(0u, TAG_ENV, step, 0u, cfg.env_k0 + (uint32_t)e, cfg.env_k1);
  uint64_t o;
  double rew;
  M::template step<true, double>(m, s, a, u01(p.w[0]), u01(p.w[2]), u01(p.w[3]), sp, o, rew);
  ec.total += ec.disc * rew;
  ec.disc *= m.gamma;
  ec.steps += 1;
  ec.last_a = a;
  ec.last_o = o;
  ep_store_state<State>(ec, sp);
  if (!(ec.disc > cfg.sim_eps) || M::terminal(m, sp) || ec.steps >= cfg.max_steps) {
    ec.active = 0;
    return;
  }
  TreeCtl& tc = trees[e];
  const uint32_t slot = tc.root * (uint32_t)m.nA + (uint32_t)a;
  if (cfg.updater == POMCP_BELIEF_EXTERNAL) {
    ec.update_pending = 1;  // SIR kernels follow; the tree is discarded (src/solver.jl:278-281)
    tc.root_N0 = 0;
    return;
  }
  const uint32_t child = find_child<PID>(P, m, slot, o);
  if (cfg.updater == POMCP_BELIEF_EXACT2) {
    if (child != 0) {  // src/updater.jl:37: the child with its subtree
      tc.root = child;
      tc.root_N0 = P.node_N[child];
    } else {           // src/updater.jl:31-35: a new node carrying update(node_belief_updater, b_old.B, a, o)
      const double nb = bayes2<PID>(m, P.node_bel[tc.root], a, o);
      const unsigned long long id = atomicAdd(&P.ctr->node_count, 1ull);
      if (id >= P.max_nodes) { ec.status = POMCP_ERR_POOL_EXHAUSTED; ec.active = 0; return; }
      P.node_bel[id] = nb;
      tc.root = (uint32_t)id;
      tc.root_N0 = 0;
    }
    return;
  }
  // RootUpdater with the DeadReinvigorator (src/updater.jl:13-27, src/particle_filter.jl:85-101)
  if (child == 0 || P.node_N[child] == 0) {
    ec.status = POMCP_ERR_PARTICLE_DEPLETION;
    ec.active = 0;
    return;
  }
  tc.root = child;
  tc.root_N0 = P.node_N[child];
  tc.n_bel = 0;  // counted by k_ep_gather
  ec.update_pending = 1;
}

// unweighted update: the new root's particle collection = the states pushed at it, gathered from the episode's
// segment of the particle log (order is not preserved: rand(b) is uniform over the collection)
template <class State>
__global__ void __launch_bounds__(256) k_ep_gather(const __grid_constant__ Pool P, TreeCtl* trees, EpisodeCtl* eps,
                                                   State* bel_buf, long long cap) {
  const int e = blockIdx.x;
  if (!eps[e].update_pending) return;
  TreeCtl& tc = trees[e];
  __shared__ unsigned int s_count;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  const unsigned long long n_log = min(*tc.log_ctr, tc.log_cap);
  const uint32_t root = tc.root;
  State* out = bel_buf + (size_t)e * (size_t)cap;
  const State* log_state = (const State*)P.log_state;
  for (unsigned long long i = threadIdx.x; i < n_log; i += blockDim.x) {
    const unsigned long long li = tc.log_base + i;
    if (P.log_node[li] == root) {
      const unsigned int k = atomicAdd(&s_count, 1u);
      if ((long long)k < cap) out[k] = log_state[li];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    tc.bel = out;
    tc.n_bel = min((long long)s_count, cap);
    if (s_count == 0) {  // an empty collection is a depleted belief (src/particle_filter.jl:85-92)
      eps[e].status = POMCP_ERR_PARTICLE_DEPLETION;
      eps[e].active = 0;
    }
  }
}

// SIR mode: every tree is thrown away after its decision, so the pool is recycled: wipe what was used and put
// the roots back at nodes 0 .. n - 1
__global__ void k_ep_pool_wipe(const __grid_constant__ Pool P, int nA, int nO, uint32_t init_N, double init_V) {
  const unsigned long long n_nodes = min(P.ctr->node_count, (unsigned long long)P.max_nodes);
  const unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  const unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  const double v0 = (init_N > 0 && fabs(init_V) != INFINITY) ? init_V * (double)init_N : 0.0;
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
__global__ void k_ep_pool_restart(const __grid_constant__ Pool P, TreeCtl* trees, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e == 0) { P.ctr->node_count = (unsigned long long)n; }
  if (e < n) { trees[e].root = (uint32_t)e; trees[e].root_N0 = 0; }
}

// The SIR source of an episode: the action and the observation come from the episode's control block, and an
// episode with no update pending skips the kernel.
template <int PID>
struct EpisodeSrc {
  using M = ModelT<PID>;
  using State = typename M::State;
  const State* states;
  DevModel m;
  const EpisodeCtl* ep;
  uint32_t k0, k1;
  static constexpr int CEIL_EXP = ModelSrc<PID>::CEIL_EXP;
  static constexpr int NCLASS = ModelSrc<PID>::NCLASS;
  static constexpr bool CLASS_IN = ModelSrc<PID>::CLASS_IN;
  __device__ __forceinline__ bool identity() const { return src().identity(); }
  __device__ __forceinline__ int wclass(const State& sp) const { return src().wclass(sp); }
  __device__ __forceinline__ double class_weight(int c) const { return src().class_weight(c); }
  __device__ __forceinline__ bool skip() const { return ep->update_pending == 0; }
  __device__ __forceinline__ ModelSrc<PID> src() const {
    ModelSrc<PID> s;
    s.states = states; s.m = m; s.a = ep->last_a; s.obs = ep->last_o; s.k0 = k0; s.k1 = k1;
    return s;
  }
  __device__ __forceinline__ bool eval(long long i, State& sp, double& w) const { return src().eval(i, sp, w); }
  __device__ __forceinline__ bool eval_from(long long i, const State& s, State& sp, double& w) const {
    return src().eval_from(i, s, sp, w);
  }
  __device__ __forceinline__ void propagate(long long i, State& sp) const { src().propagate(i, sp); }
  __device__ __forceinline__ State load(long long i) const { return pf_ld_state(states + i); }
  __device__ __forceinline__ void propagate_from(long long i, const State& s, State& sp) const {
    src().propagate_from(i, s, sp);
  }
};

// after the SIR kernels of a step: the resampled set is the belief; an update with no weight left ends the episode
__global__ void k_ep_sir_done(TreeCtl* trees, EpisodeCtl* eps, const PfCtl* const* ctls, const void* new_buf,
                              long long sir_n, int state_bytes, int n) {
  const int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= n || !eps[e].update_pending) return;
  const int st = (int)ctls[e]->status;
  if (st != POMCP_OK) { eps[e].status = st; eps[e].active = 0; return; }
  trees[e].bel = (const char*)new_buf + (size_t)e * (size_t)sir_n * (size_t)state_bytes;
  trees[e].n_bel = sir_n;
}

}  // namespace pomcp

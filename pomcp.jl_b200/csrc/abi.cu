// libpomcp_b200.so — host side of the C ABI declared in include/pomcp_b200.h.
// One handle = one device + one stream + one tree.  All kernels are in
// tree.cuh (search) and pf.cuh (particle filter); this file owns memory, launch schedules,
// status mapping and the NCCL root merge.  There is no CPU fallback anywhere.
#include <cuda_runtime.h>
#include <dlfcn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/pomcp_b200.h"
#include "episode.cuh"
#include "pf.cuh"
#include "reroot.cuh"
#include "tree.cuh"

using namespace pomcp;

namespace {

thread_local std::string g_create_error;

struct PhaseTimer {
  std::vector<cudaEvent_t> ev;  // pairs
  int used = 0;
  long long launches = 0;
  void reset() { used = 0; launches = 0; }
  cudaEvent_t next() {
    if (used == (int)ev.size()) {
      cudaEvent_t e;
      cudaEventCreate(&e);
      ev.push_back(e);
    }
    return ev[used++];
  }
  double total_ms() {
    double t = 0;
    for (int i = 0; i + 1 < used; i += 2) {
      float ms = 0;
      if (cudaEventElapsedTime(&ms, ev[i], ev[i + 1]) == cudaSuccess) t += ms;
    }
    return t;
  }
  void destroy() {
    for (auto e : ev) cudaEventDestroy(e);
    ev.clear();
  }
};

// ---- NCCL through dlopen: the library itself has no link-time NCCL dependency
typedef struct { char internal[128]; } nccl_uid;
typedef void* nccl_comm;
struct NcclApi {
  void* lib = nullptr;
  int (*GetUniqueId)(nccl_uid*) = nullptr;
  int (*CommInitRank)(nccl_comm*, int, nccl_uid, int) = nullptr;
  int (*AllReduce)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t) = nullptr;
  int (*CommDestroy)(nccl_comm) = nullptr;
  const char* (*GetErrorString)(int) = nullptr;
  bool load(std::string& err) {
    if (lib) return true;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* n : names) {
      lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
      if (lib) break;
    }
    if (!lib) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
    GetUniqueId = (int (*)(nccl_uid*))dlsym(lib, "ncclGetUniqueId");
    CommInitRank = (int (*)(nccl_comm*, int, nccl_uid, int))dlsym(lib, "ncclCommInitRank");
    AllReduce = (int (*)(const void*, void*, size_t, int, int, nccl_comm, cudaStream_t))dlsym(lib, "ncclAllReduce");
    CommDestroy = (int (*)(nccl_comm))dlsym(lib, "ncclCommDestroy");
    GetErrorString = (const char* (*)(int))dlsym(lib, "ncclGetErrorString");
    if (!GetUniqueId || !CommInitRank || !AllReduce || !CommDestroy) { err = "NCCL symbols missing"; return false; }
    return true;
  }
};
NcclApi g_nccl;
constexpr int NCCL_FLOAT64 = 8, NCCL_SUM = 0;

}  // namespace

struct pomcp_handle {
  int device = 0;
  cudaStream_t stream = nullptr;
  pomcp_problem prob{};
  pomcp_solver_params sp{};
  DevModel dm{};
  double* d_eta = nullptr;
  uint32_t* d_eta_thr = nullptr;  // RockSample sensor thresholds, compact [(y*n+x)*k+rock]
  int eta_n = 0;
  int nA = 0, nO = 0, state_bytes = 4, levels = 0;
  bool hashed = false, exact = false, unweighted = true;
  double log_ratio = 0;
  double* d_gpow = nullptr;
  Pool pool{};
  size_t ht_entries = 0, ev_entries = 0;
  PathBuf pb{};
  int n_slots = 0, inflight_min = 8, inflight_max = 1024, ramp_div = 32, rollouts = 32, sm_count = 148;
  bool inflight_min_auto = true;  // no explicit inflight_min: max(8, tree_queries / 512) per search
  void* d_bel[2] = {nullptr, nullptr};
  size_t bel_cap[2] = {0, 0};
  int bel_cur = 0, bel_kind = 0;  // 0 none, 1 initial distribution, 2 particles, 3 exact two-state belief (EXACT2)
  bool exact2 = false;            // belief_mode = EXACT2: nodes carry P(s = 1)
  double root_p = 0.0;            // EXACT2: belief of a fresh root (written to node_bel by fresh_root)
  double* d_scalar = nullptr;     // one double of scratch
  long long n_bel = 0;
  uint32_t root = 0, epoch = 0;  // root / root_visits: tree 0's (the only tree unless n_trees > 1)
  long long root_visits = 0;
  int n_trees = 1;               // trees searched side by side in one launch (root parallelism inside the GPU)
  int res_index = 0;             // slot of d_res that holds the decision of the search in flight: a tree's own (0),
                                 // the merge over this device's trees (n_trees) or over the ranks (n_trees + 1)
  std::vector<uint32_t> roots;
  std::vector<long long> tree_visits;
  TreeCtl* d_trees = nullptr;
  TreeCtl* h_trees = nullptr;    // pinned staging copy
  unsigned long long nodes_used = 1, log_used = 0;
  SearchResult* d_res = nullptr;
  SearchResult* h_res = nullptr;
  DevCtr* h_ctr = nullptr;
  PfCtl* d_pfctl = nullptr;
  PfCtl* h_pfctl = nullptr;
  NodeStats* d_ns = nullptr;
  NodeStats* h_ns = nullptr;
  unsigned long long* d_cum = nullptr;
  size_t cum_cap = 0;
  // one zero-initialised control block per scan / resample: [0..7] PfCtl, [8] ticket or barrier, [9..] tile status
  unsigned long long* d_scan_status = nullptr;
  unsigned long long* d_pf_regions = nullptr;  // fused SIR: two control blocks, each zeroed by the launch that uses the other
  int pf_flip = 0;
  size_t scan_cap = 0;
  void* d_tmp = nullptr;
  size_t tmp_cap = 0;
  void* d_tmp2 = nullptr;
  size_t tmp2_cap = 0;
  float4* d_flush = nullptr;
  size_t flush_n = 0;
  PhaseTimer timers[POMCP_PHASE_COUNT];
  cudaEvent_t user_ev[8] = {};
  pomcp_counters host_ctr{};
  long long launches = 0, last_waves = 0;
  bool pending = false;
  uint32_t* d_newid = nullptr;
  size_t newid_cap = 0;
  double* d_state_values = nullptr;  // FOValue table
  long long n_state_values = 0;
  void* d_init = nullptr;  // particles drawn from the initial distribution (reinvigorator)
  size_t init_cap = 0;
  bool compact_always = false;  // POMCP_COMPACT_ALWAYS=1: compact the pool at every re-root (tests)
  long long compactions = 0;
  bool inflight_max_auto = true;    // inflight_max was left to the library: it is also capped relative to the budget
  bool force_streaming_pf = false;  // POMCP_PF_STREAMING=1: always use the three-kernel SIR path
  bool pf_legacy_streaming = false;  // POMCP_PF_LEGACY_STREAMING=1: the max-scaled (look-back) kernels even where the tiled ones apply
  bool pf_two_barriers = false;     // POMCP_PF_TWO_BARRIERS=1: never use the one-barrier (a-priori scale) kernel
  bool pf_nocoop = false;           // POMCP_PF_NOCOOP=1: plain launch of the one-barrier kernel (developer timing)
  long long pf_retries = 0;         // one-barrier launches that fell back to the max-scaled kernels
  // device-resident episodes
  EpisodeCtl* d_eps = nullptr;
  unsigned long long* d_ep_logctr = nullptr;
  void* d_ep_bel[2] = {nullptr, nullptr};
  size_t ep_bel_cap[2] = {0, 0};
  nccl_comm comm = nullptr;
  int nranks = 1;
  std::string err;
};

namespace {

#define CK(call)                                                                                     \
  do {                                                                                               \
    cudaError_t e__ = (call);                                                                        \
    if (e__ != cudaSuccess) {                                                                        \
      h->err = std::string(#call) + ": " + cudaGetErrorString(e__);                                  \
      return POMCP_ERR_CUDA;                                                                         \
    }                                                                                                \
  } while (0)

#define DISPATCH_PID(pid, ...)                                                       \
  switch (pid) {                                                                     \
    case POMCP_TIGER: { constexpr int PID = POMCP_TIGER; __VA_ARGS__; } break;       \
    case POMCP_BABY: { constexpr int PID = POMCP_BABY; __VA_ARGS__; } break;         \
    case POMCP_ROCKSAMPLE: { constexpr int PID = POMCP_ROCKSAMPLE; __VA_ARGS__; } break; \
    default: { constexpr int PID = POMCP_LIGHTDARK1D; __VA_ARGS__; } break;          \
  }

int fail(pomcp_handle* h, int st, const std::string& msg) {
  if (h) h->err = msg; else g_create_error = msg;
  return st;
}

int model_dims(const pomcp_problem* p, int& nA, int& nO, int& sb) {
  switch (p->problem_id) {
    case POMCP_TIGER: nA = 3; nO = 2; sb = 4; return POMCP_OK;
    case POMCP_BABY: nA = 2; nO = 2; sb = 4; return POMCP_OK;
    case POMCP_ROCKSAMPLE:
      if (p->rs_n < 1 || p->rs_n > 15 || p->rs_k < 0 || p->rs_k > POMCP_MAX_ROCKS) return POMCP_ERR_INVALID_ARG;
      nA = 5 + p->rs_k; nO = 3; sb = 4; return POMCP_OK;
    case POMCP_LIGHTDARK1D: nA = 3; nO = 0; sb = 16; return POMCP_OK;
  }
  return POMCP_ERR_INVALID_ARG;
}

template <class T>
int dev_alloc(pomcp_handle* h, T** p, size_t n) {
  CK(cudaMalloc((void**)p, std::max<size_t>(n, 1) * sizeof(T)));
  return POMCP_OK;
}

int ensure_bytes(pomcp_handle* h, void** p, size_t* cap, size_t need) {
  if (need <= *cap) return POMCP_OK;
  if (*p) CK(cudaFree(*p));
  *p = nullptr;
  size_t nc = std::max<size_t>(need, 256);
  CK(cudaMalloc(p, nc));
  *cap = nc;
  return POMCP_OK;
}

int launch_check(pomcp_handle* h, const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    h->err = std::string(what) + ": " + cudaGetErrorString(e);
    return POMCP_ERR_CUDA;
  }
  h->launches++;
  return POMCP_OK;
}

int pool_fill(pomcp_handle* h, unsigned long long n_nodes) {
  if (n_nodes == 0) return POMCP_OK;
  unsigned long long work = n_nodes * (unsigned long long)h->nA * (unsigned long long)std::max(h->nO, 1);
  int blocks = (int)std::min<unsigned long long>((work + 255) / 256, 148ull * 16);
  k_pool_fill<<<blocks, 256, 0, h->stream>>>(h->pool, h->nA, h->nO, n_nodes, (uint32_t)h->sp.init_N, h->sp.init_V,
                                             h->exact ? 1 : 0);
  return launch_check(h, "k_pool_fill");
}

// RootNode(0, b, {}) (src/solver.jl:278-281): wipe what the previous tree used, root = node 0
int fresh_root(pomcp_handle* h) {
  int rc = pool_fill(h, std::min<unsigned long long>(h->nodes_used, h->pool.max_nodes));
  if (rc) return rc;
  if (h->hashed) CK(cudaMemsetAsync(h->pool.ht, 0, h->ht_entries * sizeof(uint32_t), h->stream));
  if (h->pool.ev_key) CK(cudaMemsetAsync(h->pool.ev_key, 0, h->ev_entries * sizeof(unsigned long long), h->stream));
  // keep cumulative counters: only the allocation cursors restart
  CK(cudaMemcpyAsync(h->h_ctr, h->pool.ctr, sizeof(DevCtr), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->h_ctr->node_count = (unsigned long long)h->n_trees;  // roots are nodes 0 .. n_trees - 1
  h->h_ctr->log_count = 0;
  h->h_ctr->pool_overflow = h->h_ctr->log_overflow = h->h_ctr->replay_overflow = 0;
  CK(cudaMemcpyAsync(h->pool.ctr, h->h_ctr, sizeof(DevCtr), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->exact2 && h->bel_kind == 3) {
    for (int t = 0; t < h->n_trees; ++t) k_set_node_bel<<<1, 32, 0, h->stream>>>(h->pool.node_bel, (uint32_t)t, nullptr, h->root_p);
    rc = launch_check(h, "k_set_node_bel");
    if (rc) return rc;
  }
  h->root = 0;
  h->root_visits = 0;
  for (int t = 0; t < h->n_trees; ++t) { h->roots[(size_t)t] = (uint32_t)t; h->tree_visits[(size_t)t] = 0; }
  h->nodes_used = (unsigned long long)h->n_trees;
  h->log_used = 0;
  return POMCP_OK;
}

SearchCfg make_cfg(pomcp_handle* h) {
  SearchCfg c{};
  c.c = h->sp.c;
  c.init_V = h->sp.init_V;
  c.init_N = h->sp.init_N;
  c.log_ratio = h->log_ratio;
  c.est_value = h->sp.estimator_value;
  c.state_values = h->d_state_values;
  c.n_state_values = h->n_state_values;
  c.max_depth = h->sp.max_depth;
  c.gpow = h->d_gpow;
  c.levels = h->levels;
  c.belief_unweighted = h->unweighted ? 1 : 0;
  c.estimator = h->sp.estimator;
  c.key0 = (uint32_t)h->sp.seed;
  c.key1 = (uint32_t)(h->sp.seed >> 32);
  c.rank = (uint32_t)(h->sp.rank * h->n_trees);  // stream id of tree t: rank * n_trees + t
  c.epoch = h->epoch;
  c.trees = h->d_trees;
  c.n_trees = h->n_trees;
  c.bel = (h->bel_kind == 2) ? h->d_bel[h->bel_cur] : nullptr;
  c.n_bel = h->n_bel;
  c.root = h->root;
  c.root_N0 = (uint32_t)h->root_visits;
  c.dpw = h->sp.dpw_observation ? 1 : 0;
  c.dpa = h->sp.dpw_action ? 1 : 0;
  c.dpw_solver = (h->sp.dpw_solver || h->sp.dpw_observation || h->sp.dpw_action) ? 1 : 0;
  c.exact2 = (h->exact2 && h->bel_kind == 3) ? 1 : 0;
  c.k_act = h->sp.k_action;
  c.alpha_act = h->sp.alpha_action;
  c.k_obs = h->sp.k_observation;
  c.alpha_obs = h->sp.alpha_observation;
  c.rollouts = h->rollouts;
  // Warps per worker: one per 32 rollouts.  33..64 rollouts also run on ONE warp, two trajectories per lane (1.3x the
  // time of 32, tree.cuh), and that is what the library does when the caller asks for more simulations in flight
  // than two-warp teams can keep co-resident (8 per SM at 16 warps per SM): the request can only be met by workers
  // of one warp, 16 per SM.
  c.team_warps = h->exact ? 1 : (h->rollouts > 64 ? 4 : (h->rollouts + 31) / 32);
  if (c.team_warps == 2 && h->inflight_max > 8 * h->sm_count) c.team_warps = 1;
  c.inflight_min = h->inflight_min;
  c.inflight_max = h->inflight_max;
  c.ramp_div = h->ramp_div;
  return c;
}

int reset_search_flags(pomcp_handle* h) {
  // query / worker tickets, any_nonterminal and the overflow flags are per search
  size_t off = offsetof(DevCtr, next_query);
  CK(cudaMemsetAsync((char*)h->pool.ctr + off, 0, sizeof(DevCtr) - off, h->stream));
  return POMCP_OK;
}

// One persistent launch of the batched workers: four-warp CTAs, 4 / team_warps workers per CTA, workers dealt to
// the cfg.n_trees trees round-robin, admission ramped on the device.  The per-tree table (cfg.trees) must be in
// place on the stream.
// Admission floor = tree_queries / ADMISSION_FLOOR_DIV simulations in flight from the first query on (see below)
#ifndef ADMISSION_FLOOR_DIV
#define ADMISSION_FLOOR_DIV 512
#endif
int launch_batched_workers(pomcp_handle* h, SearchCfg cfg, long long Q) {
  const int nteams = 4 / cfg.team_warps;
  size_t path_bytes = ((size_t)std::max(h->levels, 1) * (16 + (size_t)h->state_bytes) + 15) & ~(size_t)15;
  cfg.path_smem = path_bytes * nteams <= 40 * 1024 ? 1 : 0;
  cfg.path_bytes = (int)path_bytes;
  size_t smem = cfg.path_smem ? path_bytes * nteams : 0;
  cfg.eta_off = (int)smem;
  cfg.eta_n = h->eta_n;
  cfg.eta_thr = h->d_eta_thr;
  smem += (size_t)h->eta_n * sizeof(uint32_t);
  const int T = std::max(cfg.n_trees, 1);
  // Cap relative to the budget (when the caller left inflight_max to the library): no more than Q / 256 simulations of
  // a tree in flight (at least 64).  Measured on RockSample(7,8), 292 paired episodes against the sequential oracle at
  // 1e5 queries: 195 in flight -0.07 +- 0.14, 390 (this rule) +0.31 +- 0.16, 781 -0.32 +- 0.16, 2368 -0.81 +- 0.18 of
  // 15.7; at 1e6 queries the full 2368 cost nothing (+0.10 +- 0.14).  DESIGN.md 4.1.
  // The rule is applied where it was measured to matter: on Baby and Tiger 195 / 390 / 1184 simulations in flight give
  // the same root values and actions at 1e5 queries (12 seeds against the oracle) while the search takes 8.1 / 4.6 / 2.7
  // ms, and LightDark's plain POMCP tree is one level deep - those problems keep their 8 teams per SM.
  long long cap = h->inflight_max;
  if (h->inflight_max_auto && h->prob.problem_id == POMCP_ROCKSAMPLE)
    cap = std::min<long long>(cap, std::max<long long>(64, Q / 256) * T);
  long long workers = std::min<long long>(cap, std::max<long long>(Q, 1) * T);
  int blocks = (int)((workers + nteams - 1) / nteams);
  DISPATCH_PID(h->prob.problem_id, {
    // no more workers than can be co-resident: a worker that is not resident does not search
    auto launch = [&](auto kern) {
      int per_sm = 0, sms = 0;
      cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, 128, smem);
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
      if (per_sm > 0 && sms > 0) blocks = std::min(blocks, per_sm * sms);
      // workers are dealt to the trees round-robin: the in-flight cap, the admission floor and the ramp are per tree
      cfg.inflight_max = (int)std::max<long long>(1, std::min<long long>(cap, (long long)blocks * nteams) / T);
      // Admission floor relative to the budget: the simulations started while the tree holds fewer than
      // Q / 512 visits are two thousandths of the search, so running that many at once costs the final tree
      // nothing and removes most of the ramp from a large search.  Measured (scripts/probe.py knob --knob inflight_min,
      // RockSample(7,8), 1e6 queries, 12 seeds): floor 976 / 1953 / 2368: V - V(oracle) -1.63 +- 0.15 / -1.43 +- 0.17 /
      // -1.78 +- 0.16, the oracle's action 12/12 each, 8.73 / 8.42 / 8.42 ms; at 1e5 queries (floor 97 / 195, cap 390)
      // 292 paired episodes against the oracle: -0.09 +- 0.15 / +0.24 +- 0.17.  (Q / 1024 until the end of round 2.)
      if (h->inflight_min_auto)
        cfg.inflight_min = (int)std::min<long long>(std::max<long long>(8, Q / ADMISSION_FLOOR_DIV), cfg.inflight_max);
      cfg.inflight_min = std::min(cfg.inflight_min, cfg.inflight_max);
      h->last_waves = cfg.inflight_max;  // reported as pomcp_counters::waves: the admission cap in effect, per tree
      kern<<<blocks, 128, smem, h->stream>>>(h->pool, h->dm, cfg, h->pb, Q);
    };
    // action widening and (on discrete observations) observation widening are separate instantiations: the POMCP path
    // pays nothing for them.  Continuous observations always carry the observation-widening code (run-time flag).
    constexpr bool H = ModelT<PID>::HASHED;
    const bool dpo = !H && cfg.dpw;
    auto pick = [&](auto dpa_c, auto dpo_c) {
      constexpr bool A = decltype(dpa_c)::value, O = decltype(dpo_c)::value;
      if (cfg.team_warps == 4) launch(k_search_workers<PID, 4, 1, A, O>);
      else if (cfg.team_warps == 2) launch(k_search_workers<PID, 2, 2, A, O>);
      else launch(k_search_workers<PID, 1, 4, A, O>);
    };
    using T_ = std::true_type;
    using F_ = std::false_type;
    using H_ = std::integral_constant<bool, H>;
    if (cfg.dpa) { if (dpo) pick(T_{}, T_{}); else pick(T_{}, H_{}); }
    else { if (dpo) pick(F_{}, T_{}); else pick(F_{}, H_{}); }
  });
  return launch_check(h, "k_search_workers");
}

int finish_search(pomcp_handle* h);

int enqueue_search(pomcp_handle* h, long long Q, const double* d_uni, long long n_uni, const double* d_rets,
                   long long n_rets) {
  if (h->bel_kind == 0) return fail(h, POMCP_ERR_INVALID_ARG, "no belief set");
  if (h->sp.estimator == POMCP_EST_STATE_VALUE && !h->d_state_values)
    return fail(h, POMCP_ERR_INVALID_ARG, "POMCP_EST_STATE_VALUE needs pomcp_set_state_values first");
  if (Q < 0) Q = h->sp.tree_queries;
  if (Q >= (1ll << 32)) return fail(h, POMCP_ERR_INVALID_ARG, "tree_queries must be < 2^32");
  int rc;
  if (h->pending) {  // the staging copy of the tree table is re-used: drain the previous search first
    rc = finish_search(h);
    if (rc != POMCP_OK && rc != POMCP_ERR_ALL_SAMPLES_TERMINAL && rc != POMCP_ERR_NAN_VALUE) return rc;
  }
  rc = reset_search_flags(h);
  if (rc) return rc;
  SearchCfg cfg = make_cfg(h);
  h->roots[0] = h->root;
  h->tree_visits[0] = h->root_visits;
  for (int t = 0; t < h->n_trees; ++t) {
    TreeCtl& tc = h->h_trees[t];
    tc.next_query = 0;
    tc.bel = cfg.bel;
    tc.n_bel = cfg.n_bel;
    tc.root = h->roots[(size_t)t];
    tc.root_N0 = (uint32_t)h->tree_visits[(size_t)t];
    tc.key0 = cfg.key0; tc.key1 = cfg.key1;
    tc.rank = cfg.rank + (uint32_t)t;
    tc.epoch = cfg.epoch;
    tc.any_nonterminal = 0;
    tc._pad = 0;
    tc.log_ctr = &h->pool.ctr->log_count;  // device address: one log, one cursor (several trees share both)
    tc.log_base = 0;
    tc.log_cap = h->pool.log_cap;
  }
  CK(cudaMemcpyAsync(h->d_trees, h->h_trees, sizeof(TreeCtl) * (size_t)h->n_trees, cudaMemcpyHostToDevice, h->stream));
  const bool replay = d_uni != nullptr;
  cfg.uni = d_uni; cfg.n_uni = n_uni; cfg.rets = d_rets; cfg.n_rets = n_rets;
  PhaseTimer& ts = h->timers[POMCP_PHASE_SEARCH];
  PhaseTimer& tr = h->timers[POMCP_PHASE_ROLLOUT];
  ts.reset(); tr.reset();
  h->last_waves = 0;
  CK(cudaEventRecord(ts.next(), h->stream));
  if (h->exact || replay) {
    if (!h->exact) return fail(h, POMCP_ERR_INVALID_ARG, "replay needs search_mode = exact");
    DISPATCH_PID(h->prob.problem_id, {
      if (replay) k_search_exact<PID, true><<<1, 32, 0, h->stream>>>(h->pool, h->dm, cfg, h->pb, Q, h->d_res);
      else k_search_exact<PID, false><<<1, 32, 0, h->stream>>>(h->pool, h->dm, cfg, h->pb, Q, h->d_res);
    });
    rc = launch_check(h, "k_search_exact");
    if (rc) return rc;
    ts.launches++;
  } else {
    CK(cudaEventRecord(tr.next(), h->stream));
    rc = launch_batched_workers(h, cfg, Q);
    if (rc) return rc;
    CK(cudaEventRecord(tr.next(), h->stream));
    tr.launches++;
    ts.launches++;
  }
  h->root_visits += Q;
  k_root_result<<<h->n_trees, 32, 0, h->stream>>>(h->pool, h->nA, h->d_trees, h->exact ? 1 : 0, h->sp.init_V, h->d_res);
  rc = launch_check(h, "k_root_result");
  if (rc) return rc;
  ts.launches++;
  if (h->n_trees > 1) {  // merge the trees of this device like K10 merges the devices (src/solver.jl:60-70)
    k_trees_merge<<<1, 32, 0, h->stream>>>(h->n_trees, h->nA, h->sp.init_V, h->d_res);
    rc = launch_check(h, "k_trees_merge");
    if (rc) return rc;
    ts.launches++;
  }
  CK(cudaEventRecord(ts.next(), h->stream));
  h->epoch++;
  h->pending = true;
  h->res_index = h->n_trees > 1 ? h->n_trees : 0;
  return POMCP_OK;
}

// wait for the stream, pull result + counters back
int finish_search(pomcp_handle* h) {
  // h_res[0] = the decision (the merged result when several trees ran), h_res[1 + t] = tree t's own
  const int T = h->n_trees;
  CK(cudaMemcpyAsync(h->h_res, h->d_res + h->res_index, sizeof(SearchResult), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_res + 1, h->d_res, sizeof(SearchResult) * (size_t)T, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_ctr, h->pool.ctr, sizeof(DevCtr), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->pending = false;
  h->nodes_used = std::max<unsigned long long>(h->h_ctr->node_count, 1);
  h->log_used = std::min<unsigned long long>(h->h_ctr->log_count, h->pool.log_cap);
  for (int t = 0; t < T; ++t) h->tree_visits[(size_t)t] = h->h_res[1 + t].root_N;
  h->root_visits = h->tree_visits[0];
  if (h->h_res->status == POMCP_ERR_POOL_EXHAUSTED)
    h->err = "node pool or particle log exhausted (raise max_nodes / max_log)";
  return h->h_res->status;
}

void copy_out(pomcp_handle* h, int32_t* action_out, int64_t* N, double* V, int32_t nA) {
  if (action_out) *action_out = h->h_res->action;
  for (int a = 0; a < h->nA && a < nA; ++a) {
    if (N) N[a] = h->h_res->N[a];
    if (V) V[a] = h->h_res->V[a];
  }
}

template <int PID>
ModelSrc<PID> make_src(pomcp_handle* h, const void* states, int a, uint64_t obs, uint64_t seed) {
  ModelSrc<PID> s;
  s.states = (const typename ModelT<PID>::State*)states;
  s.m = h->dm;
  s.a = a;
  s.obs = obs;
  s.k0 = (uint32_t)seed;
  s.k1 = (uint32_t)(seed >> 32);
  return s;
}

int ensure_scan(pomcp_handle* h, long long n) {
  static_assert(sizeof(PfCtl) <= 64, "PfCtl must fit the first 8 words of the control block");
  size_t tiles = (size_t)((n + SCAN_TILE - 1) / SCAN_TILE) + 12;
  int rc = ensure_bytes(h, (void**)&h->d_scan_status, &h->scan_cap, std::max<size_t>(tiles, 512) * sizeof(unsigned long long));
  if (rc) return rc;
  h->d_pfctl = (PfCtl*)h->d_scan_status;
  CK(cudaMemsetAsync(h->d_scan_status, 0, tiles * sizeof(unsigned long long), h->stream));
  return POMCP_OK;
}

// K6 -> K7 -> K8 on an arbitrary weight source.  `allow_ceil`: try the one-barrier kernel with the a-priori
// scale first (sources with a weight ceiling); it reports PF_RETRY_MAX_SCALE in the control block when that scale
// does not fit, and the caller (run_resample_sync) then runs the max-scaled kernels.
template <class Src, class Out>
int run_resample(pomcp_handle* h, const Src& src, long long n, uint32_t ru, long long n_out, Out* d_out, bool allow_ceil) {
  int rc;
  const bool ceil_ok = allow_ceil && Src::CEIL_EXP != PF_NO_CEIL && !h->pf_two_barriers;
  constexpr bool blocked = sizeof(typename Src::State) <= 8;
  // fused single-launch path when the whole set fits one co-resident grid
  static int max_grid = -1;  // per <Src,Out> instantiation
  if (max_grid < 0) {
    int per_sm = 0, per_sm_c = 0, sms = 0;
    if (blocked) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pf_fused<Src, Out>, PF_THREADS, 0);
    else cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_pf_fused_striped<Src, Out>, PF_THREADS, 0);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm_c, k_pf_fused_ceil<Src, Out, blocked>, PF_THREADS, 0);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, h->device);
    max_grid = std::min(std::min(per_sm, per_sm_c) * sms, PF_MAX_GRID);
  }
  // split the set evenly over the co-resident grid: `rows` consecutive particles per thread (<= PF_ITEMS)
  long long rows_ll = max_grid > 0 ? (n + (long long)max_grid * PF_THREADS - 1) / ((long long)max_grid * PF_THREADS) : 99;
  int rows = (int)std::max<long long>(rows_ll, 1);
  long long grid = (n + (long long)PF_THREADS * rows - 1) / ((long long)PF_THREADS * rows);
  if (rows <= PF_ITEMS && grid <= max_grid && !h->force_streaming_pf) {
    // The control block (PfCtl, barrier counter, CTA totals) must start zeroed.  There are two; a launch uses
    // one and zeroes the other for the next launch, so no memset sits between the caller and the kernel.
    if (!h->d_pf_regions) {
      CK(cudaMalloc((void**)&h->d_pf_regions, 2 * PF_CTRL_WORDS * sizeof(unsigned long long)));
      CK(cudaMemsetAsync(h->d_pf_regions, 0, 2 * PF_CTRL_WORDS * sizeof(unsigned long long), h->stream));
    }
    unsigned long long* region = h->d_pf_regions + (size_t)h->pf_flip * PF_CTRL_WORDS;
    unsigned long long* other = h->d_pf_regions + (size_t)(1 - h->pf_flip) * PF_CTRL_WORDS;
    h->pf_flip ^= 1;
    h->d_pfctl = (PfCtl*)region;
    PfCtl* ctl = h->d_pfctl;
    unsigned long long* cta_tot = region + 9;
    unsigned int* bar = (unsigned int*)(region + 8);
    Src src_copy = src;
    void* args[] = {(void*)&src_copy, (void*)&n, (void*)&ctl, (void*)&ru, (void*)&n_out, (void*)&d_out,
                    (void*)&cta_tot, (void*)&bar, (void*)&rows, (void*)&other};
    if (ceil_ok) {
      // grid <= what the occupancy calculator says is co-resident; the cooperative launch makes the driver
      // guarantee it (POMCP_PF_NOCOOP=1: plain launch, for timing the difference)
      if (h->pf_nocoop) {
        k_pf_fused_ceil<Src, Out, blocked><<<(unsigned)grid, PF_THREADS, 0, h->stream>>>(src_copy, n, ctl, ru, n_out, d_out,
                                                                                         cta_tot, bar, rows, other);
      } else {
        CK(cudaLaunchCooperativeKernel((const void*)k_pf_fused_ceil<Src, Out, blocked>, dim3((unsigned)grid),
                                       dim3(PF_THREADS), args, 0, h->stream));
      }
    } else if (blocked) {
      CK(cudaLaunchCooperativeKernel((const void*)k_pf_fused<Src, Out>, dim3((unsigned)grid), dim3(PF_THREADS), args,
                                     0, h->stream));
    } else {  // wide states: striped variant; its control block is zeroed here (it does not reset the other one)
      CK(cudaMemsetAsync(other, 0, PF_CTRL_WORDS * sizeof(unsigned long long), h->stream));
      CK(cudaLaunchCooperativeKernel((const void*)k_pf_fused_striped<Src, Out>, dim3((unsigned)grid), dim3(PF_THREADS),
                                     args, 0, h->stream));
    }
    rc = launch_check(h, "k_pf_fused");
    if (rc) return rc;
    h->timers[POMCP_PHASE_PF].launches += 1;
    return POMCP_OK;
  }
  // per-particle weights between the passes (the weight-class sources of the tiled path look them up again instead)
  if (!(ceil_ok && !h->pf_legacy_streaming && Src::NCLASS > 0)) {
    rc = ensure_bytes(h, (void**)&h->d_cum, &h->cum_cap, (size_t)n * sizeof(unsigned long long));
    if (rc) return rc;
  }
  rc = ensure_scan(h, n);
  if (rc) return rc;
  if (ceil_ok && !h->pf_legacy_streaming) {
    // tiled streaming path at the a-priori scale: weigh + quantise + tile sums -> tile prefix -> scan + emit per tile
    const long long ntiles = (n + PF2_TILE - 1) / PF2_TILE;
    unsigned long long* tile_sum = h->d_scan_status + 10;  // 16-byte aligned
    // CTAs walk several tiles when the model has a weight-class table (built once per CTA)
    const unsigned grid = (unsigned)std::min<long long>(ntiles, Src::NCLASS > 0 ? (long long)h->sm_count * 8 : ntiles);
    k_pf_weigh_q<Src><<<grid, PF2_THREADS, 0, h->stream>>>(src, n, ntiles, (unsigned long long*)h->d_cum, tile_sum,
                                                           h->d_pfctl);
    rc = launch_check(h, "k_pf_weigh_q");
    if (rc) return rc;
    // weight-class sources: a CTA owns a contiguous chunk of tiles and the emit kernel (same grid) derives offsets,
    // total and status from the per-CTA sums; the others: one sum per tile and the single-CTA prefix kernel
    if (Src::NCLASS == 0) {
      k_pf_tile_prefix<<<1, 1024, 0, h->stream>>>(tile_sum, ntiles, h->d_pfctl);
      rc = launch_check(h, "k_pf_tile_prefix");
      if (rc) return rc;
    }
    k_pf_emit_tiled<Src, Out><<<grid, PF2_THREADS, 0, h->stream>>>(src, n, ntiles, (const unsigned long long*)h->d_cum,
                                                                   tile_sum, h->d_pfctl, ru, n_out, d_out);
    rc = launch_check(h, "k_pf_emit_tiled");
    if (rc) return rc;
    h->timers[POMCP_PHASE_PF].launches += Src::NCLASS == 0 ? 3 : 2;
    return POMCP_OK;
  }
  // max-scaled streaming path: weigh -> total(s) -> scan + emit (picks the scale by the same rule)
  int blocks = (int)std::min<long long>((n + 1023) / 1024, 148 * 8);
  k_pf_weigh<Src><<<blocks, 256, 0, h->stream>>>(src, n, (double*)h->d_cum, h->d_pfctl);
  rc = launch_check(h, "k_pf_weigh");
  if (rc) return rc;
  k_pf_total<<<(int)std::min<long long>((n + 2047) / 2048, 148 * 8), 256, 0, h->stream>>>((const double*)h->d_cum, n,
                                                                                         h->d_pfctl, Src::CEIL_EXP);
  rc = launch_check(h, "k_pf_total");
  if (rc) return rc;
  int tiles = (int)((n + PF2_TILE - 1) / PF2_TILE);
  k_pf_scan_emit<Src, Out><<<tiles, PF2_THREADS, 0, h->stream>>>(src, n, (const double*)h->d_cum, h->d_pfctl, ru, n_out,
                                                                 d_out, h->d_scan_status + 9, h->d_scan_status + 8);
  rc = launch_check(h, "k_pf_scan_emit");
  if (rc) return rc;
  h->timers[POMCP_PHASE_PF].launches += 3;
  return POMCP_OK;
}

// run_resample + the status read-back every caller needs; falls back to the max-scaled kernels when the
// one-barrier kernel asks for it.  Leaves the stream synchronised and the status in h->h_pfctl->status.
template <class Src, class Out>
int run_resample_sync(pomcp_handle* h, const Src& src, long long n, uint32_t ru, long long n_out, Out* d_out,
                      cudaEvent_t done_event = nullptr) {
  int rc = run_resample(h, src, n, ru, n_out, d_out, true);
  if (rc) return rc;
  if (done_event) CK(cudaEventRecord(done_event, h->stream));
  CK(cudaMemcpyAsync(h->h_pfctl, h->d_pfctl, sizeof(PfCtl), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  if (h->h_pfctl->status != PF_RETRY_MAX_SCALE) return POMCP_OK;
  h->pf_retries++;
  rc = run_resample(h, src, n, ru, n_out, d_out, false);
  if (rc) return rc;
  if (done_event) CK(cudaEventRecord(done_event, h->stream));  // the phase timer then covers both attempts
  CK(cudaMemcpyAsync(h->h_pfctl, h->d_pfctl, sizeof(PfCtl), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return POMCP_OK;
}

// Pool compaction after a re-root (see reroot.cuh): keeps the sub-tree under h->root, which becomes node 0.
int compact_pool(pomcp_handle* h) {
  const long long n_nodes = (long long)std::min<unsigned long long>(h->nodes_used, h->pool.max_nodes);
  const long long n_log = (long long)h->log_used;
  const int nA = h->nA, nO = h->nO, sb = h->state_bytes;
  Pool& P = h->pool;
  int rc = ensure_bytes(h, (void**)&h->d_newid, &h->newid_cap, (size_t)n_nodes * sizeof(uint32_t));
  if (rc) return rc;
  // 1. mark + rank
  rc = ensure_scan(h, n_nodes);
  if (rc) return rc;
  int tiles = (int)((n_nodes + SCAN_TILE - 1) / SCAN_TILE);
  k_rank_subtree<<<tiles, SCAN_THREADS, 0, h->stream>>>(P.node_pslot, h->root, (uint32_t)nA, n_nodes, h->d_newid,
                                                        h->d_scan_status + 9, h->d_scan_status + 8);
  rc = launch_check(h, "k_rank_subtree");
  if (rc) return rc;
  k_scan_total<<<1, 32, 0, h->stream>>>(h->d_scan_status + 9, tiles, &h->d_pfctl->total);
  rc = launch_check(h, "k_scan_total");
  if (rc) return rc;
  // the new root becomes node 0 (its rank is swapped with the rank-0 node's; the scratch word is PfCtl::_pad's neighbour)
  uint32_t* d_root_rank = (uint32_t*)&h->d_pfctl->total_ceil;
  k_root_rank<<<1, 32, 0, h->stream>>>(h->d_newid, h->root, d_root_rank);
  k_root_first<<<(int)((n_nodes + 255) / 256), 256, 0, h->stream>>>(h->d_newid, n_nodes, h->root, d_root_rank);
  rc = launch_check(h, "k_root_first");
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->h_pfctl, h->d_pfctl, sizeof(PfCtl), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const long long kept = (long long)h->h_pfctl->total;
  if (kept <= 0 || kept > n_nodes) return fail(h, POMCP_ERR_CUDA, "pool compaction: bad sub-tree size");
  // 2. copy the kept nodes through a scratch block
  auto al = [](size_t x) { return (x + 255) & ~(size_t)255; };
  size_t o_N = 0, o_fl = o_N + al((size_t)kept * 4), o_ps = o_fl + al((size_t)kept * 4),
         o_obs = o_ps + al((size_t)kept * 4), o_aN = o_obs + al((h->hashed || h->exact2) ? (size_t)kept * 8 : 0),
         o_aM = o_aN + al((size_t)kept * nA * 4), o_aV = o_aM + al((size_t)kept * nA * 4),
         o_ch = o_aV + al((size_t)kept * nA * 8), o_am = o_ch + al((size_t)kept * nA * nO * 4),
         total = o_am + al(P.node_amask ? (size_t)kept * 4 : 0);
  rc = ensure_bytes(h, &h->d_tmp, &h->tmp_cap, total);
  if (rc) return rc;
  char* base = (char*)h->d_tmp;
  PoolScratch T{(uint32_t*)(base + o_N), (uint32_t*)(base + o_fl), (uint32_t*)(base + o_ps),
                (unsigned long long*)(base + o_obs), (uint32_t*)(base + o_aN), (uint32_t*)(base + o_aM),
                (double*)(base + o_aV), (uint32_t*)(base + o_ch), (uint32_t*)(base + o_am)};
  k_copy_kept<<<(int)((n_nodes + 255) / 256), 256, 0, h->stream>>>(P, T, h->d_newid, h->root, nA, nO, n_nodes);
  rc = launch_check(h, "k_copy_kept");
  if (rc) return rc;
  auto d2d = [&](void* dst, const void* src, size_t bytes) {
    return bytes ? cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, h->stream) : cudaSuccess;
  };
  CK(d2d(P.node_N, T.node_N, (size_t)kept * 4));
  CK(d2d(P.node_flags, T.node_flags, (size_t)kept * 4));
  CK(d2d(P.node_pslot, T.node_pslot, (size_t)kept * 4));
  if (h->hashed) CK(d2d(P.node_obs, T.node_obs, (size_t)kept * 8));
  if (h->exact2) CK(d2d(P.node_bel, T.node_obs, (size_t)kept * 8));  // the scratch column is shared: never both
  CK(d2d(P.act_N, T.act_N, (size_t)kept * nA * 4));
  CK(d2d(P.act_M, T.act_M, (size_t)kept * nA * 4));
  CK(d2d(P.act_V, T.act_V, (size_t)kept * nA * 8));
  if (!h->hashed) CK(d2d(P.child, T.child, (size_t)kept * nA * nO * 4));
  if (P.node_amask) CK(d2d(P.node_amask, T.node_amask, (size_t)kept * 4));
  if (n_nodes > kept) {
    unsigned long long cnt = (unsigned long long)(n_nodes - kept);
    int blocks = (int)std::min<unsigned long long>((cnt * nA * std::max(nO, 1) + 255) / 256, 148ull * 16);
    k_pool_fill_range<<<blocks, 256, 0, h->stream>>>(P, nA, nO, (unsigned long long)kept, cnt, (uint32_t)h->sp.init_N,
                                                     h->sp.init_V, h->exact ? 1 : 0);
    rc = launch_check(h, "k_pool_fill_range");
    if (rc) return rc;
  }
  if (h->hashed) {
    CK(cudaMemsetAsync(P.ht, 0, h->ht_entries * sizeof(uint32_t), h->stream));
    k_rehash<<<(int)((kept + 255) / 256), 256, 0, h->stream>>>(P, kept);
    rc = launch_check(h, "k_rehash");
    if (rc) return rc;
  }
  // 3. the particle log
  long long kept_log = 0;
  if (n_log > 0) {
    size_t o_st = al((size_t)n_log * 4);
    rc = ensure_bytes(h, &h->d_tmp2, &h->tmp2_cap, o_st + (size_t)n_log * sb);
    if (rc) return rc;
    rc = ensure_scan(h, n_log);
    if (rc) return rc;
    int ltiles = (int)((n_log + SCAN_TILE - 1) / SCAN_TILE);
    uint32_t* out_node = (uint32_t*)h->d_tmp2;
    void* out_state = (char*)h->d_tmp2 + o_st;
    if (sb == 4)
      k_log_compact<uint32_t><<<ltiles, SCAN_THREADS, 0, h->stream>>>(P.log_node, (const uint32_t*)P.log_state,
                                                                     h->d_newid, n_log, out_node, (uint32_t*)out_state,
                                                                     h->d_scan_status + 9, h->d_scan_status + 8);
    else
      k_log_compact<LDState><<<ltiles, SCAN_THREADS, 0, h->stream>>>(P.log_node, (const LDState*)P.log_state,
                                                                    h->d_newid, n_log, out_node, (LDState*)out_state,
                                                                    h->d_scan_status + 9, h->d_scan_status + 8);
    rc = launch_check(h, "k_log_compact");
    if (rc) return rc;
    k_scan_total<<<1, 32, 0, h->stream>>>(h->d_scan_status + 9, ltiles, &h->d_pfctl->total);
    rc = launch_check(h, "k_scan_total");
    if (rc) return rc;
    CK(cudaMemcpyAsync(h->h_pfctl, h->d_pfctl, sizeof(PfCtl), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    kept_log = (long long)h->h_pfctl->total;
    CK(d2d(P.log_node, out_node, (size_t)kept_log * 4));
    CK(d2d(P.log_state, out_state, (size_t)kept_log * sb));
  }
  k_set_counts<<<1, 32, 0, h->stream>>>(P.ctr, (unsigned long long)kept, (unsigned long long)kept_log);
  rc = launch_check(h, "k_set_counts");
  if (rc) return rc;
  CK(cudaStreamSynchronize(h->stream));
  h->root = 0;
  h->nodes_used = (unsigned long long)kept;
  h->log_used = (unsigned long long)kept_log;
  h->compactions++;
  return POMCP_OK;
}

}  // namespace

// ===================================================================== C ABI
extern "C" {

int32_t pomcp_abi_version(void) { return POMCP_B200_ABI_VERSION; }

const char* pomcp_status_string(int32_t s) {
  switch (s) {
    case POMCP_OK: return "OK";
    case POMCP_ERR_ALL_SAMPLES_TERMINAL: return "ALL_SAMPLES_TERMINAL";
    case POMCP_ERR_PARTICLE_DEPLETION: return "PARTICLE_DEPLETION";
    case POMCP_ERR_POOL_EXHAUSTED: return "POOL_EXHAUSTED";
    case POMCP_ERR_UNSUPPORTED: return "UNSUPPORTED";
    case POMCP_ERR_INVALID_ARG: return "INVALID_ARG";
    case POMCP_ERR_CUDA: return "CUDA_ERROR";
    case POMCP_ERR_NCCL: return "NCCL_ERROR";
    case POMCP_ERR_NAN_VALUE: return "NAN_VALUE";
  }
  return "UNKNOWN";
}

const char* pomcp_last_error(const pomcp_handle* h) { return h ? h->err.c_str() : g_create_error.c_str(); }

void pomcp_default_solver_params(pomcp_solver_params* o) {  // src/constructor.jl:8-18
  std::memset(o, 0, sizeof(*o));
  o->eps = 0.01;
  o->max_depth = INT64_MAX;
  o->c = 1.0;
  o->tree_queries = 100;
  o->seed = 0;
  o->init_V = 0.0;
  o->init_N = 0;
  o->num_sparse_actions = 0;
  o->estimator = POMCP_EST_RANDOM_ROLLOUT;
  o->belief_mode = POMCP_BELIEF_UNWEIGHTED;
  o->search_mode = POMCP_MODE_BATCHED;
  o->dpw_observation = 0;
  o->k_observation = 10.0;  /* src/constructor.jl:48-49 */
  o->alpha_observation = 0.5;
  o->dpw_action = 0;
  o->k_action = 10.0;  /* src/constructor.jl:50-51 */
  o->alpha_action = 0.5;
}

void pomcp_default_problem(int32_t pid, pomcp_problem* p) {
  std::memset(p, 0, sizeof(*p));
  p->problem_id = pid;
  p->tiger_r_listen = -1; p->tiger_r_findtiger = -100; p->tiger_r_escapetiger = 10; p->tiger_p_listen_correctly = 0.85;
  p->baby_r_feed = -5; p->baby_r_hungry = -10; p->baby_p_become_hungry = 0.1; p->baby_p_cry_when_hungry = 0.8;
  p->baby_p_cry_when_not_hungry = 0.1;
  p->rs_n = 7; p->rs_k = 0; p->rs_start_x = 0; p->rs_start_y = 3; p->rs_half_eff_dist = 20;
  p->rs_r_good = 10; p->rs_r_bad = -10; p->rs_r_exit = 10;
  p->ld_correct_r = 10; p->ld_incorrect_r = -10; p->ld_init_mean = 2; p->ld_init_std = 3;
  p->discount = (pid == POMCP_TIGER || pid == POMCP_ROCKSAMPLE) ? 0.95 : 0.9;
}

int32_t pomcp_num_actions(const pomcp_problem* p) {
  int nA, nO, sb;
  return model_dims(p, nA, nO, sb) == POMCP_OK ? nA : -1;
}
int32_t pomcp_state_bytes(const pomcp_problem* p) {
  int nA, nO, sb;
  return model_dims(p, nA, nO, sb) == POMCP_OK ? sb : -1;
}

int32_t pomcp_destroy(pomcp_handle* h) {
  if (!h) return POMCP_OK;
  cudaSetDevice(h->device);
  if (h->stream) cudaStreamSynchronize(h->stream);
  if (h->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(h->comm);
  void* ptrs[] = {h->pool.node_amask, h->pool.act_nc, h->pool.act_ng, h->pool.ev_key, h->pool.ev_node, h->pool.ev_state,
                  h->d_eta, h->d_eta_thr, h->d_gpow, h->pool.node_N, h->pool.node_flags, h->pool.act_N, h->pool.act_M,
                  h->pool.act_V, h->pool.child, h->pool.ht, h->pool.node_pslot, h->pool.node_obs, h->pool.log_node,
                  h->pool.log_state, h->pool.ctr, h->pb.path_slot, h->pb.path_node, h->pb.path_r, h->pb.path_state,
                  h->d_bel[0], h->d_bel[1], h->d_res, h->d_ns,
                  h->d_cum, h->d_scan_status, h->d_pf_regions, h->d_tmp, h->d_tmp2, h->d_flush, h->d_newid, h->d_init, h->d_state_values,
                  h->d_trees, h->pool.node_bel, h->d_scalar, h->d_eps, h->d_ep_logctr, h->d_ep_bel[0], h->d_ep_bel[1]};
  for (void* p : ptrs)
    if (p) cudaFree(p);
  if (h->h_res) cudaFreeHost(h->h_res);
  if (h->h_trees) cudaFreeHost(h->h_trees);
  if (h->h_ctr) cudaFreeHost(h->h_ctr);
  if (h->h_pfctl) cudaFreeHost(h->h_pfctl);
  if (h->h_ns) cudaFreeHost(h->h_ns);
  for (auto& t : h->timers) t.destroy();
  for (auto e : h->user_ev)
    if (e) cudaEventDestroy(e);
  if (h->stream) cudaStreamDestroy(h->stream);
  delete h;
  return POMCP_OK;
}

int32_t pomcp_create(const pomcp_problem* p, const pomcp_solver_params* sp, int32_t device, pomcp_handle** out) {
  if (!p || !sp || !out) return fail(nullptr, POMCP_ERR_INVALID_ARG, "null argument");
  *out = nullptr;
  int nA, nO, sb;
  if (model_dims(p, nA, nO, sb) != POMCP_OK) return fail(nullptr, POMCP_ERR_INVALID_ARG, "bad problem descriptor");
  if (!(p->discount > 0.0) || !(p->discount < 1.0))
    return fail(nullptr, POMCP_ERR_INVALID_ARG, "discount must be in (0,1) (gamma == 1 is broken upstream, src/solver.jl:97-102)");
  if (!(sp->eps > 0.0) || !(sp->eps < 1.0)) return fail(nullptr, POMCP_ERR_INVALID_ARG, "eps must be in (0,1)");
  if (sp->num_sparse_actions > 0)
    return fail(nullptr, POMCP_ERR_UNSUPPORTED, "num_sparse_actions > 0 (random action sub-sampling) is not supported");
  const bool fwc = sp->estimator == POMCP_EST_FWC_ROLLOUT || sp->estimator == POMCP_EST_FO_FWC_ROLLOUT;
  if (sp->estimator != POMCP_EST_RANDOM_ROLLOUT && sp->estimator != POMCP_EST_CONSTANT &&
      !(fwc && p->problem_id == POMCP_BABY) &&
      !(sp->estimator == POMCP_EST_STATE_VALUE && p->problem_id != POMCP_LIGHTDARK1D))
    return fail(nullptr, POMCP_ERR_UNSUPPORTED,
                "estimator (FeedWhenCrying rollouts exist for BabyPOMDP only, state-value tables for discrete states only)");
  if (sp->init_N < 0 || sp->max_depth < 0) return fail(nullptr, POMCP_ERR_INVALID_ARG, "init_N/max_depth < 0");
  if (sp->dpw_action && (!(sp->k_action >= 0.0) || !(sp->alpha_action >= 0.0)))
    return fail(nullptr, POMCP_ERR_INVALID_ARG, "k_action / alpha_action < 0");
  if (sp->dpw_observation) {
    if (sp->belief_mode != POMCP_BELIEF_EXTERNAL)
      return fail(nullptr, POMCP_ERR_UNSUPPORTED, "observation widening runs with an external belief updater");
    if (!(sp->k_observation >= 0.0) || !(sp->alpha_observation >= 0.0))
      return fail(nullptr, POMCP_ERR_INVALID_ARG, "k_observation / alpha_observation < 0");
  }
  const int T = sp->n_trees > 1 ? sp->n_trees : 1;
  if (T > POMCP_MAX_TREES) return fail(nullptr, POMCP_ERR_INVALID_ARG, "n_trees > POMCP_MAX_TREES");
  if (T > 1 && sp->search_mode == POMCP_MODE_EXACT)
    return fail(nullptr, POMCP_ERR_INVALID_ARG, "n_trees > 1 needs search_mode = batched (exact mode is one sequential tree)");
  if (sp->rank < 0 || ((long long)sp->rank + 1) * T > 256)
    return fail(nullptr, POMCP_ERR_INVALID_ARG, "(rank + 1) * n_trees must be <= 256 (8-bit Philox stream id)");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail(nullptr, POMCP_ERR_CUDA, "no CUDA device (there is no CPU fallback)");
  if (device < 0 || device >= ndev) return fail(nullptr, POMCP_ERR_INVALID_ARG, "bad device index");
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10)
    return fail(nullptr, POMCP_ERR_CUDA, "device is not sm_100 class (library is built for sm_100a only)");

  pomcp_handle* h = new pomcp_handle();
  auto bail = [&](int st) {
    g_create_error = h->err;
    pomcp_destroy(h);
    return st;
  };
#define CKC(expr) do { int rc__ = (expr); if (rc__) return bail(rc__); } while (0)
#define CKU(call) do { cudaError_t e__ = (call); if (e__ != cudaSuccess) { h->err = std::string(#call) + ": " + cudaGetErrorString(e__); return bail(POMCP_ERR_CUDA); } } while (0)
  h->device = device;
  h->prob = *p;
  h->sp = *sp;
  h->nA = nA; h->nO = nO; h->state_bytes = sb;
  h->hashed = (nO == 0);
  h->exact = sp->search_mode == POMCP_MODE_EXACT;
  h->n_trees = T;
  h->roots.assign((size_t)T, 0u);
  h->tree_visits.assign((size_t)T, 0);
  h->unweighted = sp->belief_mode == POMCP_BELIEF_UNWEIGHTED;
  h->exact2 = sp->belief_mode == POMCP_BELIEF_EXACT2;
  if (h->exact2 && p->problem_id != POMCP_TIGER && p->problem_id != POMCP_BABY) {
    h->err = "belief_mode EXACT2 (the problem's exact discrete updater as node_belief_updater) exists for the two-state problems";
    return bail(POMCP_ERR_UNSUPPORTED);
  }
  if (h->exact2 && sp->dpw_action) { h->err = "EXACT2 with action widening is not supported"; return bail(POMCP_ERR_UNSUPPORTED); }
  h->force_streaming_pf = std::getenv("POMCP_PF_STREAMING") != nullptr;
  h->pf_legacy_streaming = std::getenv("POMCP_PF_LEGACY_STREAMING") != nullptr;
  h->pf_two_barriers = std::getenv("POMCP_PF_TWO_BARRIERS") != nullptr;
  h->pf_nocoop = std::getenv("POMCP_PF_NOCOOP") != nullptr;
  h->compact_always = std::getenv("POMCP_COMPACT_ALWAYS") != nullptr;
  CKU(cudaSetDevice(device));
  CKU(cudaStreamCreateWithFlags(&h->stream, cudaStreamNonBlocking));

  // horizon tables: gamma^depth < eps cut-off (src/solver.jl:82) and steps_to_eps (src/solver.jl:100)
  h->log_ratio = std::log(sp->eps) / std::log(p->discount);
  long long d = 0;
  while (std::pow(p->discount, (double)d) >= sp->eps && d <= 0xffff) ++d;
  if (d > 0xfff0) { h->err = "horizon log(eps)/log(discount) too long (> 65520 levels)"; return bail(POMCP_ERR_INVALID_ARG); }
  h->levels = (int)std::min<long long>(d, sp->max_depth);
  std::vector<double> gp((size_t)d + 1);
  for (long long i = 0; i <= d; ++i) gp[(size_t)i] = std::pow(p->discount, (double)i);
  CKC(dev_alloc(h, &h->d_gpow, gp.size()));
  CKU(cudaMemcpy(h->d_gpow, gp.data(), gp.size() * sizeof(double), cudaMemcpyHostToDevice));

  // device model block
  DevModel& m = h->dm;
  m.pid = p->problem_id; m.nA = nA; m.nO = nO; m.state_bytes = sb; m.gamma = p->discount;
  m.t_r_listen = p->tiger_r_listen; m.t_r_find = p->tiger_r_findtiger; m.t_r_escape = p->tiger_r_escapetiger;
  m.t_p_correct = p->tiger_p_listen_correctly;
  m.b_r_feed = p->baby_r_feed; m.b_r_hungry = p->baby_r_hungry; m.b_p_hungry = p->baby_p_become_hungry;
  m.b_p_cry_h = p->baby_p_cry_when_hungry; m.b_p_cry_nh = p->baby_p_cry_when_not_hungry;
  m.b_thr_hungry = (unsigned long long)std::ceil(std::min(std::max(p->baby_p_become_hungry, 0.0), 1.0) * 4294967296.0);
  m.rs_n = p->rs_n; m.rs_k = p->rs_k;
  m.rs_start = (unsigned)p->rs_start_x | ((unsigned)p->rs_start_y << 4);
  m.rs_r_good = p->rs_r_good; m.rs_r_bad = p->rs_r_bad; m.rs_r_exit = p->rs_r_exit;
  m.rs_rtab[0] = 0.0; m.rs_rtab[1] = p->rs_r_exit; m.rs_rtab[2] = p->rs_r_bad; m.rs_rtab[3] = p->rs_r_good;
  m.ld_correct = p->ld_correct_r; m.ld_incorrect = p->ld_incorrect_r; m.ld_mean = p->ld_init_mean; m.ld_std = p->ld_init_std;
  std::memset(m.rock_at, -1, sizeof(m.rock_at));
  if (p->problem_id == POMCP_ROCKSAMPLE) {
    std::vector<double> eta(256 * POMCP_MAX_ROCKS, 0.0);
    for (int i = 0; i < p->rs_k; ++i) {
      if (p->rs_rock_x[i] < 0 || p->rs_rock_x[i] >= p->rs_n || p->rs_rock_y[i] < 0 || p->rs_rock_y[i] >= p->rs_n) {
        h->err = "rock outside the grid";
        return bail(POMCP_ERR_INVALID_ARG);
      }
      m.rock_at[p->rs_rock_y[i] * 16 + p->rs_rock_x[i]] = (signed char)i;
    }
    if (p->rs_start_x < 0 || p->rs_start_x >= p->rs_n || p->rs_start_y < 0 || p->rs_start_y >= p->rs_n) {
      h->err = "start outside the grid";
      return bail(POMCP_ERR_INVALID_ARG);
    }
    for (int y = 0; y < p->rs_n; ++y)
      for (int x = 0; x < p->rs_n; ++x)
        for (int i = 0; i < p->rs_k; ++i) {
          double dx = (double)(x - p->rs_rock_x[i]), dy = (double)(y - p->rs_rock_y[i]);
          double dist = std::sqrt(dx * dx + dy * dy);
          eta[(size_t)(y * 16 + x) * POMCP_MAX_ROCKS + i] = 0.5 * (1.0 + std::exp2(-dist / p->rs_half_eff_dist));
        }
    CKC(dev_alloc(h, &h->d_eta, eta.size()));
    CKU(cudaMemcpy(h->d_eta, eta.data(), eta.size() * sizeof(double), cudaMemcpyHostToDevice));
    m.rs_eta = h->d_eta;
    // the same event as u < eta for u = word * 2^-32:  word <= ceil(eta * 2^32) - 1  (eta >= 0.5)
    h->eta_n = p->rs_n * p->rs_n * p->rs_k;
    std::vector<uint32_t> thr((size_t)std::max(h->eta_n, 1), 0u);
    for (int y = 0; y < p->rs_n; ++y)
      for (int x = 0; x < p->rs_n; ++x)
        for (int i = 0; i < p->rs_k; ++i) {
          double e = eta[(size_t)(y * 16 + x) * POMCP_MAX_ROCKS + i];
          thr[(size_t)(y * p->rs_n + x) * p->rs_k + i] = (uint32_t)(std::ceil(e * 4294967296.0) - 1.0);
        }
    CKC(dev_alloc(h, &h->d_eta_thr, thr.size()));
    CKU(cudaMemcpy(h->d_eta_thr, thr.data(), thr.size() * sizeof(uint32_t), cudaMemcpyHostToDevice));
  }

  // simulations in flight (batched mode) and rollouts per leaf
  // default: every resident warp a tree walker on RockSample (16 per SM, 2368 on a 148-SM B200: its rollouts run two
  // trajectories per lane, so a worker of 64 rollouts per leaf is one warp), 8 two-warp teams per SM (1184) elsewhere.
  // Measured at equal wall time the wider setting gives the better decision; inflight_max = 8 per SM
  // restores the narrower one (value estimates closer to the sequential algorithm's at equal queries).
  h->sm_count = prop.multiProcessorCount;
  h->inflight_max = sp->inflight_max > 0 ? std::min(sp->inflight_max, 4096)
                                         : (p->problem_id == POMCP_ROCKSAMPLE ? 16 : 8) * prop.multiProcessorCount;
  // every tree needs a worker of its own (workers are dealt to the trees round-robin and stay with their tree): a cap
  // below n_trees would leave the trees beyond it unsearched and still merge their empty roots
  h->inflight_max = std::max(h->inflight_max, T);
  h->inflight_max_auto = sp->inflight_max <= 0;
  h->inflight_min = sp->inflight_min > 0 ? sp->inflight_min : 8;
  h->inflight_min_auto = sp->inflight_min <= 0;
  h->inflight_min = std::max(1, std::min(h->inflight_min, h->inflight_max));
  h->ramp_div = sp->inflight_ramp_div > 0 ? sp->inflight_ramp_div : 32;
  if (sp->rollouts_per_leaf < 0 || sp->rollouts_per_leaf > (h->exact ? 64 : 128)) {
    h->err = "rollouts_per_leaf must be in 1..128 (1..64 in exact mode)";
    return bail(POMCP_ERR_INVALID_ARG);
  }
  h->rollouts = sp->rollouts_per_leaf > 0 ? sp->rollouts_per_leaf : (h->exact ? 1 : 64);
  h->n_slots = h->exact ? 1 : h->inflight_max;

  // pool sizing
  size_t free_b = 0, total_b = 0;
  CKU(cudaMemGetInfo(&free_b, &total_b));
  long long Q = std::max<long long>(sp->tree_queries, 1);
  size_t per_node = 8 + (size_t)nA * 16 + (h->hashed ? 12 + 8 : (size_t)nA * nO * 4) +
                    (sp->dpw_observation ? (size_t)nA * 8 + 32 * (12 + (size_t)sb) : 0);
  // unweighted mode keeps the tree across decisions (re-rooting, no compaction yet): room for ~64 of them
  long long want = sp->max_nodes > 0 ? sp->max_nodes : ((h->unweighted || h->exact2) ? 64 * Q : 2 * Q) * T + 65536;
  long long cap_mem = (long long)((free_b * 4 / 10) / per_node);
  // node ids carry a flag in bit 31, and action slots (id * nA + a) are 32-bit as well
  const long long cap_ids = std::min<long long>((1ll << 31) - 2, ((1ll << 32) - 1) / std::max(nA, 1) - 1);
  if (sp->max_nodes > cap_ids) {
    h->err = "max_nodes * n_actions must stay below 2^32 (32-bit action slots)";
    return bail(POMCP_ERR_INVALID_ARG);
  }
  want = std::min<long long>(want, std::min<long long>(cap_mem, cap_ids));
  if (want < 16) { h->err = "not enough device memory for the node pool"; return bail(POMCP_ERR_CUDA); }
  Pool& P = h->pool;
  P.max_nodes = (uint32_t)want;
  CKC(dev_alloc(h, &P.node_N, (size_t)want));
  CKC(dev_alloc(h, &P.node_flags, (size_t)want));
  CKC(dev_alloc(h, &P.act_N, (size_t)want * nA));
  CKC(dev_alloc(h, &P.act_M, (size_t)want * nA));
  CKC(dev_alloc(h, &P.act_V, (size_t)want * nA));
  CKC(dev_alloc(h, &P.node_pslot, (size_t)want));
  if (sp->dpw_action) CKC(dev_alloc(h, &P.node_amask, (size_t)want));
  if (h->exact2) CKC(dev_alloc(h, &P.node_bel, (size_t)want));
  CKC(dev_alloc(h, &h->d_scalar, 1));
  if (h->hashed) {
    size_t e = 1;
    while (e < (size_t)want * 2) e <<= 1;
    h->ht_entries = e;
    P.ht_mask = e - 1;
    CKC(dev_alloc(h, &P.ht, e));
    CKC(dev_alloc(h, &P.node_obs, (size_t)want));
    CKU(cudaMemset(P.ht, 0, e * sizeof(uint32_t)));
  } else {
    CKC(dev_alloc(h, &P.child, (size_t)want * nA * nO));
  }
  if (sp->dpw_observation) {  // continuous or discrete observations alike (src/solver.jl:223-262)
    // one event per generated tree step: room for 8 per node (POOL_EXHAUSTED beyond)
    size_t ev = 1;
    while (ev < (size_t)want * 16) ev <<= 1;
    h->ev_entries = ev;
    P.ev_mask = ev - 1;
    CKC(dev_alloc(h, &P.act_nc, (size_t)want * nA));
    CKC(dev_alloc(h, &P.act_ng, (size_t)want * nA));
    CKC(dev_alloc(h, &P.ev_key, ev));
    CKC(dev_alloc(h, &P.ev_node, ev));
    CKU(cudaMalloc(&P.ev_state, ev * (size_t)sb));
  }
  long long log_want = 0;
  if (h->unweighted) {
    log_want = sp->max_log > 0 ? sp->max_log : want * 8;
    long long cap_log = (long long)((free_b * 2 / 10) / (4 + (size_t)sb));
    log_want = std::min(log_want, cap_log);
  }
  P.log_cap = (unsigned long long)log_want;
  CKC(dev_alloc(h, &P.log_node, (size_t)log_want));
  CKU(cudaMalloc(&P.log_state, std::max<size_t>((size_t)log_want * sb, 16)));
  CKC(dev_alloc(h, &P.ctr, 1));
  CKU(cudaMemset(P.ctr, 0, sizeof(DevCtr)));

  // per-worker path scratch
  size_t pw = (size_t)std::max(h->levels, 1) * h->n_slots;
  h->pb.stride = h->n_slots;
  CKC(dev_alloc(h, &h->pb.path_slot, pw));
  CKC(dev_alloc(h, &h->pb.path_node, pw));
  CKC(dev_alloc(h, &h->pb.path_r, pw));
  CKU(cudaMalloc(&h->pb.path_state, pw * sb));

  CKC(dev_alloc(h, &h->d_res, (size_t)T + 2));  // one per tree + the merge over the trees + the merge over the ranks
  CKC(dev_alloc(h, &h->d_ns, 1));
  CKC(dev_alloc(h, &h->d_trees, (size_t)T));
  CKU(cudaMallocHost((void**)&h->h_trees, sizeof(TreeCtl) * (size_t)T));
  CKU(cudaMallocHost((void**)&h->h_res, sizeof(SearchResult) * ((size_t)T + 1)));
  CKU(cudaMallocHost((void**)&h->h_ctr, sizeof(DevCtr)));
  CKU(cudaMallocHost((void**)&h->h_pfctl, sizeof(PfCtl)));
  CKU(cudaMallocHost((void**)&h->h_ns, sizeof(NodeStats)));
  std::memset(h->h_res, 0, sizeof(SearchResult) * ((size_t)T + 1));

  CKC(ensure_scan(h, 1));
  h->nodes_used = want;  // first fill covers the whole pool
  CKC(fresh_root(h));
  CKU(cudaStreamSynchronize(h->stream));
#undef CKC
#undef CKU
  *out = h;
  return POMCP_OK;
}

int32_t pomcp_set_belief_particles(pomcp_handle* h, const void* states, int64_t n) {
  if (!h || !states || n <= 0 || n >= (1ll << 32)) return fail(h, POMCP_ERR_INVALID_ARG, "bad particle array");
  cudaSetDevice(h->device);
  int b = h->bel_cur;
  int rc = ensure_bytes(h, &h->d_bel[b], &h->bel_cap[b], (size_t)n * h->state_bytes);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_bel[b], states, (size_t)n * h->state_bytes, cudaMemcpyHostToDevice, h->stream));
  h->n_bel = n;
  h->bel_kind = 2;
  return fresh_root(h);
}

int32_t pomcp_set_state_values(pomcp_handle* h, const double* values, int64_t n) {
  if (!h || !values || n <= 0) return fail(h, POMCP_ERR_INVALID_ARG, "bad state-value table");
  cudaSetDevice(h->device);
  if (h->d_state_values) CK(cudaFree(h->d_state_values));
  h->d_state_values = nullptr;
  CK(cudaMalloc((void**)&h->d_state_values, (size_t)n * sizeof(double)));
  CK(cudaMemcpyAsync(h->d_state_values, values, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->n_state_values = n;
  return POMCP_OK;
}

int32_t pomcp_set_belief_initial(pomcp_handle* h) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  h->bel_kind = 1;
  h->n_bel = 0;
  return fresh_root(h);
}

// initialize_belief(up, BoolDistribution(p)) for a planner whose node_belief_updater is the problem's exact
// two-state updater: RootNode(0, b, {}) with b = P(s = 1) (src/updater.jl:45-47).
int32_t pomcp_set_belief_exact2(pomcp_handle* h, double p1) {
  if (!h || !(p1 >= 0.0) || !(p1 <= 1.0)) return fail(h, POMCP_ERR_INVALID_ARG, "belief must be a probability");
  cudaSetDevice(h->device);
  if (!h->exact2) return fail(h, POMCP_ERR_UNSUPPORTED, "planner was not created with belief_mode = EXACT2");
  h->bel_kind = 3;
  h->n_bel = 0;
  h->root_p = p1;
  return fresh_root(h);
}

int32_t pomcp_get_belief_exact2(pomcp_handle* h, double* p1) {
  if (!h || !p1) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (!h->exact2 || h->bel_kind != 3) return fail(h, POMCP_ERR_INVALID_ARG, "no exact belief set");
  CK(cudaMemcpyAsync(p1, h->pool.node_bel + h->root, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return POMCP_OK;
}

// make_node for a belief that is not a BeliefNode (src/solver.jl:278-281): every action() on a particle
// collection or a distribution starts from a new RootNode - here without sending the belief again.
int32_t pomcp_reset_tree(pomcp_handle* h) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (h->bel_kind == 0) return fail(h, POMCP_ERR_INVALID_ARG, "no belief set");
  return fresh_root(h);
}

int32_t pomcp_get_belief_particles(pomcp_handle* h, void* out, int64_t cap, int64_t* n) {
  if (!h || !n) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  *n = (h->bel_kind == 2) ? h->n_bel : 0;
  if (out && cap > 0 && h->bel_kind == 2) {
    size_t cnt = (size_t)std::min<int64_t>(cap, h->n_bel);
    CK(cudaMemcpyAsync(out, h->d_bel[h->bel_cur], cnt * h->state_bytes, cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
  }
  return POMCP_OK;
}

int32_t pomcp_search_async(pomcp_handle* h, int64_t Q) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  return enqueue_search(h, Q, nullptr, 0, nullptr, 0);
}

int32_t pomcp_sync(pomcp_handle* h) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (h->pending) return finish_search(h);
  CK(cudaStreamSynchronize(h->stream));
  return POMCP_OK;
}

int32_t pomcp_search(pomcp_handle* h, int64_t Q, int32_t* action_out, int64_t* N, double* V, int32_t nA) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  int rc = enqueue_search(h, Q, nullptr, 0, nullptr, 0);
  if (rc) return rc;
  rc = finish_search(h);
  if (rc == POMCP_OK || rc == POMCP_ERR_NAN_VALUE) copy_out(h, action_out, N, V, nA);
  return rc;
}

int32_t pomcp_search_replay(pomcp_handle* h, int64_t Q, const double* uni, int64_t nu, const double* rets, int64_t nr,
                            int32_t* action_out, int64_t* N, double* V, int32_t nA, int64_t* u_used,
                            int64_t* r_used) {
  if (!h || !uni || nu < 0 || nr < 0) return fail(h, POMCP_ERR_INVALID_ARG, "bad replay arrays");
  cudaSetDevice(h->device);
  if (!h->exact) return fail(h, POMCP_ERR_INVALID_ARG, "replay needs search_mode = exact");
  int rc = ensure_bytes(h, &h->d_tmp, &h->tmp_cap, (size_t)std::max<int64_t>(nu, 1) * sizeof(double));
  if (rc) return rc;
  rc = ensure_bytes(h, &h->d_tmp2, &h->tmp2_cap, (size_t)std::max<int64_t>(nr, 1) * sizeof(double));
  if (rc) return rc;
  if (nu) CK(cudaMemcpyAsync(h->d_tmp, uni, (size_t)nu * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  if (nr) CK(cudaMemcpyAsync(h->d_tmp2, rets, (size_t)nr * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  rc = enqueue_search(h, Q, (const double*)h->d_tmp, nu, (const double*)h->d_tmp2, nr);
  if (rc) return rc;
  h->epoch--;  // replay does not consume a Philox epoch
  rc = finish_search(h);
  if (u_used) *u_used = h->h_res->ucur;
  if (r_used) *r_used = h->h_res->rcur;
  if (rc == POMCP_OK) copy_out(h, action_out, N, V, nA);
  return rc;
}

static int node_stats_impl(pomcp_handle* h, const int32_t* actions, const uint64_t* obs, int32_t depth) {
  if (depth < 0 || depth > 64) return fail(h, POMCP_ERR_INVALID_ARG, "history depth must be in [0,64]");
  NodeQuery q{};
  q.depth = depth;
  for (int d = 0; d < depth; ++d) {
    if (actions[d] < 0 || actions[d] >= h->nA) return fail(h, POMCP_ERR_INVALID_ARG, "bad action in history");
    q.actions[d] = actions[d];
    q.obs[d] = obs[d];
  }
  long long nb = (h->bel_kind == 2) ? h->n_bel : 0;
  DISPATCH_PID(h->prob.problem_id, {
    k_node_stats<PID><<<1, 32, 0, h->stream>>>(h->pool, h->dm, h->root, q, h->exact ? 1 : 0, h->sp.init_V, nb,
                                               h->unweighted ? 1 : 0, h->d_ns);
  });
  int rc = launch_check(h, "k_node_stats");
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->h_ns, h->d_ns, sizeof(NodeStats), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return POMCP_OK;
}

int32_t pomcp_get_node_stats(pomcp_handle* h, const int32_t* actions, const uint64_t* obs, int32_t depth,
                             int64_t* node_N, int64_t* n_particles, int64_t* act_N, double* act_V, int32_t nA) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  int rc = node_stats_impl(h, actions, obs, depth);
  if (rc) return rc;
  if (h->h_ns->node < 0) {
    if (node_N) *node_N = -1;
    return POMCP_OK;
  }
  if (node_N) *node_N = h->h_ns->N;
  if (n_particles) *n_particles = h->h_ns->n_particles;
  for (int a = 0; a < h->nA && a < nA; ++a) {
    if (act_N) act_N[a] = h->h_ns->act_N[a];
    if (act_V) act_V[a] = h->h_ns->act_V[a];
  }
  return POMCP_OK;
}

// Tree t's own result of the last search (before the merge over trees / ranks).
int32_t pomcp_get_tree_result(pomcp_handle* h, int32_t tree, int32_t* action_out, int64_t* root_N, int64_t* act_N,
                              double* act_V, int32_t nA) {
  if (!h || tree < 0 || tree >= h->n_trees) return fail(h, POMCP_ERR_INVALID_ARG, "bad tree index");
  if (h->pending) return fail(h, POMCP_ERR_INVALID_ARG, "a search is still in flight (pomcp_sync first)");
  const SearchResult& r = h->h_res[1 + tree];
  if (action_out) *action_out = r.action;
  if (root_N) *root_N = r.root_N;
  for (int a = 0; a < h->nA && a < nA; ++a) {
    if (act_N) act_N[a] = r.N[a];
    if (act_V) act_V[a] = r.V[a];
  }
  return r.status;
}

int32_t pomcp_get_root_stats(pomcp_handle* h, int64_t* root_N, int64_t* act_N, double* act_V, int32_t nA) {
  return pomcp_get_node_stats(h, nullptr, nullptr, 0, root_N, nullptr, act_N, act_V, nA);
}

// Gather the `cnt` states pushed at node `child` from the particle log into belief buffer `nb`
// (room for `cap` states).  Enqueued + verified; leaves the stream synchronised.
static int gather_child_particles(pomcp_handle* h, uint32_t child, long long cnt, int nb, long long cap) {
  PhaseTimer& t = h->timers[POMCP_PHASE_REROOT];
  int rc = ensure_bytes(h, &h->d_bel[nb], &h->bel_cap[nb], (size_t)std::max(cap, cnt) * h->state_bytes);
  if (rc) return rc;
  long long n_log = (long long)h->log_used;
  rc = ensure_scan(h, n_log);
  if (rc) return rc;
  int tiles = (int)((n_log + SCAN_TILE - 1) / SCAN_TILE);
  if (tiles > 0) {
    if (h->state_bytes == 4)
      k_log_gather<uint32_t><<<tiles, SCAN_THREADS, 0, h->stream>>>(h->pool.log_node, (const uint32_t*)h->pool.log_state,
                                                                  child, n_log, (uint32_t*)h->d_bel[nb], cnt,
                                                                  h->d_scan_status + 9, h->d_scan_status + 8);
    else
      k_log_gather<LDState><<<tiles, SCAN_THREADS, 0, h->stream>>>(h->pool.log_node, (const LDState*)h->pool.log_state,
                                                                 child, n_log, (LDState*)h->d_bel[nb], cnt,
                                                                 h->d_scan_status + 9, h->d_scan_status + 8);
    rc = launch_check(h, "k_log_gather");
    if (rc) return rc;
    k_scan_total<<<1, 32, 0, h->stream>>>(h->d_scan_status + 9, tiles, &h->d_pfctl->total);
    rc = launch_check(h, "k_scan_total");
    if (rc) return rc;
    t.launches += 3;
  }
  CK(cudaEventRecord(t.next(), h->stream));
  CK(cudaMemcpyAsync(h->h_pfctl, h->d_pfctl, sizeof(PfCtl), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  long long found = tiles > 0 ? (long long)h->h_pfctl->total : 0;
  if (found != cnt) {
    char buf[160];
    std::snprintf(buf, sizeof buf, "particle log holds %lld states for the new root but N = %lld", found, cnt);
    return fail(h, POMCP_ERR_POOL_EXHAUSTED, buf);
  }
  return POMCP_OK;
}

static int maybe_compact(pomcp_handle* h) {
  // reclaim the dead part of the pool / particle log once a quarter of either is used
  if (h->compact_always || h->nodes_used * 4 >= h->pool.max_nodes || h->log_used * 4 >= h->pool.log_cap)
    return compact_pool(h);
  return POMCP_OK;
}

// update(::RootUpdater{<:ParticleReinvigorator}, b_old, a, o): src/updater.jl:13-27
int32_t pomcp_update_unweighted(pomcp_handle* h, int32_t a, uint64_t obs_bits) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (!h->unweighted) return fail(h, POMCP_ERR_UNSUPPORTED, "planner was created with an external belief updater");
  if (h->n_trees > 1) return fail(h, POMCP_ERR_UNSUPPORTED, "re-rooting needs n_trees = 1 (several trees share one external belief)");
  if (a < 0 || a >= h->nA) return fail(h, POMCP_ERR_INVALID_ARG, "bad action");
  PhaseTimer& t = h->timers[POMCP_PHASE_REROOT];
  t.reset();
  CK(cudaEventRecord(t.next(), h->stream));
  int32_t acts[1] = {a};
  uint64_t obs[1] = {obs_bits};
  int rc = node_stats_impl(h, acts, obs, 1);
  if (rc) return rc;
  // missing child -> handle_unseen_observation; empty collection -> reinvigorate! (src/particle_filter.jl:85-92)
  if (h->h_ns->node < 0 || h->h_ns->N == 0) return POMCP_ERR_PARTICLE_DEPLETION;
  uint32_t child = (uint32_t)h->h_ns->node;
  long long cnt = h->h_ns->N;
  int nb = 1 - h->bel_cur;
  rc = gather_child_particles(h, child, cnt, nb, cnt);
  if (rc) return rc;
  h->bel_cur = nb;
  h->n_bel = cnt;
  h->bel_kind = 2;
  h->root = child;
  h->root_visits = cnt;
  return maybe_compact(h);
}

// update(::RootUpdater, b_old, a, o) with a user-supplied (here: the problem's exact two-state) belief updater,
// src/updater.jl:30-39: the child (a, o) becomes the root with its subtree; if the tree does not hold it, a new
// node with update(node_belief_updater, b_old.B, a, o) takes its place - this updater never depletes.
int32_t pomcp_update_exact2(pomcp_handle* h, int32_t a, uint64_t obs_bits) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (!h->exact2 || h->bel_kind != 3) return fail(h, POMCP_ERR_UNSUPPORTED, "planner holds no exact two-state belief");
  if (h->n_trees > 1) return fail(h, POMCP_ERR_UNSUPPORTED, "re-rooting needs n_trees = 1");
  if (a < 0 || a >= h->nA || obs_bits >= (uint64_t)h->nO) return fail(h, POMCP_ERR_INVALID_ARG, "bad action / observation");
  int32_t acts[1] = {a};
  uint64_t obs[1] = {obs_bits};
  int rc = node_stats_impl(h, acts, obs, 1);
  if (rc) return rc;
  if (h->h_ns->node >= 0) {
    h->root = (uint32_t)h->h_ns->node;
    h->root_visits = h->h_ns->N;
    return maybe_compact(h);
  }
  DISPATCH_PID(h->prob.problem_id, {
    k_bayes2<PID><<<1, 32, 0, h->stream>>>(h->dm, h->pool.node_bel, h->root, a, obs_bits, h->d_scalar);
  });
  rc = launch_check(h, "k_bayes2");
  if (rc) return rc;
  CK(cudaMemcpyAsync(&h->root_p, h->d_scalar, sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return fresh_root(h);  // ObsNode(o, 0, new_belief, {}): an empty tree under the new belief
}

// update(::RootUpdater{R}, b_old, a, o) with the device-side reinvigorator R = SIRReinvigorator(n_min)
// (hooks of src/particle_filter.jl:37-73, call sites src/updater.jl:14-23):
//   handle_unseen_observation: child (a,o) missing or empty -> n_min particles from the SIR posterior of the
//                              old root belief become the new belief (fresh root);
//   reinvigorate!:             fewer than n_min particles at the child -> topped up from the same posterior.
int32_t pomcp_update_reinvigorated(pomcp_handle* h, int32_t a, uint64_t obs_bits, uint64_t seed, int64_t n_min,
                                   int32_t* outcome) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (!h->unweighted) return fail(h, POMCP_ERR_UNSUPPORTED, "planner was created with an external belief updater");
  if (h->n_trees > 1) return fail(h, POMCP_ERR_UNSUPPORTED, "re-rooting needs n_trees = 1 (several trees share one external belief)");
  if (a < 0 || a >= h->nA || n_min <= 0 || n_min >= (1ll << 32)) return fail(h, POMCP_ERR_INVALID_ARG, "bad a / n_min");
  if (h->bel_kind == 0) return fail(h, POMCP_ERR_INVALID_ARG, "no belief set");
  if (outcome) *outcome = 0;
  PhaseTimer& t = h->timers[POMCP_PHASE_REROOT];
  t.reset();
  CK(cudaEventRecord(t.next(), h->stream));
  int32_t acts[1] = {a};
  uint64_t obs[1] = {obs_bits};
  int rc = node_stats_impl(h, acts, obs, 1);
  if (rc) return rc;
  const bool have = h->h_ns->node >= 0 && h->h_ns->N > 0;
  const long long cnt = have ? h->h_ns->N : 0;
  const uint32_t child = have ? (uint32_t)h->h_ns->node : 0u;
  if (have && cnt >= n_min) return pomcp_update_unweighted(h, a, obs_bits);
  const long long extra = n_min - cnt;
  // the old root belief as a particle array
  const void* old = h->d_bel[h->bel_cur];
  long long n_old = h->n_bel;
  if (h->bel_kind == 1) {
    rc = ensure_bytes(h, &h->d_init, &h->init_cap, (size_t)n_min * h->state_bytes);
    if (rc) return rc;
    DISPATCH_PID(h->prob.problem_id, {
      k_init_particles<PID><<<(int)((n_min + 255) / 256), 256, 0, h->stream>>>(
          h->dm, (typename ModelT<PID>::State*)h->d_init, n_min, (uint32_t)seed, (uint32_t)(seed >> 32));
    });
    rc = launch_check(h, "k_init_particles");
    if (rc) return rc;
    old = h->d_init;
    n_old = n_min;
  }
  const int nb = 1 - h->bel_cur;
  if (have) {
    rc = gather_child_particles(h, child, cnt, nb, n_min);
    if (rc) return rc;
  } else {
    rc = ensure_bytes(h, &h->d_bel[nb], &h->bel_cap[nb], (size_t)n_min * h->state_bytes);
    if (rc) return rc;
  }
  // SIR posterior of the old belief: `extra` particles appended behind the gathered ones
  Philox4 pr = philox4x32_10(0u, TAG_RESAMPLE, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  DISPATCH_PID(h->prob.problem_id, {
    auto src = make_src<PID>(h, old, a, obs_bits, seed);
    rc = run_resample_sync(h, src, n_old, pr.w[0], extra, (typename ModelT<PID>::State*)h->d_bel[nb] + cnt);
  });
  if (rc) return rc;
  const bool drew = h->h_pfctl->status == POMCP_OK;
  if (!drew && !have) return POMCP_ERR_PARTICLE_DEPLETION;  // nothing pushed by the planner and no posterior mass
  h->bel_cur = nb;
  h->bel_kind = 2;
  if (have) {
    h->n_bel = cnt + (drew ? extra : 0);
    h->root = child;
    h->root_visits = cnt;
    if (outcome) *outcome = drew ? 1 : 0;
    return maybe_compact(h);
  }
  h->n_bel = extra;
  if (outcome) *outcome = 2;
  return fresh_root(h);
}

// update(::SIRParticleFilter, b, a, o): propagate, weight, systematic resample (test/runtests.jl:57)
int32_t pomcp_pf_update_sir(pomcp_handle* h, int32_t a, uint64_t obs_bits, uint64_t seed, int64_t n_out) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (h->bel_kind != 2) return fail(h, POMCP_ERR_INVALID_ARG, "SIR update needs a particle belief");
  if (a < 0 || a >= h->nA || n_out <= 0 || n_out >= (1ll << 32)) return fail(h, POMCP_ERR_INVALID_ARG, "bad a / n_out");
  const long long n = h->n_bel;
  int nb = 1 - h->bel_cur;
  int rc = ensure_bytes(h, &h->d_bel[nb], &h->bel_cap[nb], (size_t)n_out * h->state_bytes);
  if (rc) return rc;
  Philox4 pr = philox4x32_10(0u, TAG_RESAMPLE, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
  PhaseTimer& t = h->timers[POMCP_PHASE_PF];
  t.reset();
  CK(cudaEventRecord(t.next(), h->stream));
  cudaEvent_t t_done = t.next();
  DISPATCH_PID(h->prob.problem_id, {
    auto src = make_src<PID>(h, h->d_bel[h->bel_cur], a, obs_bits, seed);
    rc = run_resample_sync(h, src, n, pr.w[0], n_out, (typename ModelT<PID>::State*)h->d_bel[nb], t_done);
  });
  if (rc) return rc;
  if (h->h_pfctl->status != POMCP_OK) return h->h_pfctl->status;
  h->bel_cur = nb;
  h->n_bel = n_out;
  return fresh_root(h);  // a ParticleCollection is not a BeliefNode: new RootNode (src/solver.jl:278-281)
}

int32_t pomcp_resample_systematic(pomcp_handle* h, const double* w, int64_t n, double r, int64_t n_out, int32_t mode,
                                  int64_t* idx_out) {
  if (!h || !w || !idx_out || n <= 0 || n_out <= 0 || n >= (1ll << 32) || n_out >= (1ll << 32) || !(r >= 0.0) ||
      !(r < 1.0))
    return fail(h, POMCP_ERR_INVALID_ARG, "bad resample arguments");
  cudaSetDevice(h->device);
  int rc = ensure_bytes(h, &h->d_tmp, &h->tmp_cap, (size_t)n * sizeof(double));
  if (rc) return rc;
  rc = ensure_bytes(h, &h->d_tmp2, &h->tmp2_cap, (size_t)n_out * sizeof(long long));
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_tmp, w, (size_t)n * sizeof(double), cudaMemcpyHostToDevice, h->stream));
  CK(cudaMemsetAsync(h->d_tmp2, 0xff, (size_t)n_out * sizeof(long long), h->stream));
  if (mode == 0) {
    ArraySrc src{(const double*)h->d_tmp};
    rc = run_resample(h, src, n, (uint32_t)(r * 4294967296.0), n_out, (long long*)h->d_tmp2, false);
    if (rc) return rc;
  } else {
    rc = ensure_scan(h, 1);
    if (rc) return rc;
    k_resample_seq_f64<<<1, 32, 0, h->stream>>>((const double*)h->d_tmp, n, r, n_out, (long long*)h->d_tmp2, h->d_pfctl);
    rc = launch_check(h, "k_resample_seq_f64");
    if (rc) return rc;
  }
  CK(cudaMemcpyAsync(h->h_pfctl, h->d_pfctl, sizeof(PfCtl), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(idx_out, h->d_tmp2, (size_t)n_out * sizeof(long long), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return h->h_pfctl->status;
}

int32_t pomcp_rollout_batch(pomcp_handle* h, const void* states, int64_t n, int32_t steps, uint64_t seed,
                            double* ret, int32_t* steps_out) {
  if (!h || !states || !ret || n <= 0 || n >= (1ll << 32) || steps < 0) return fail(h, POMCP_ERR_INVALID_ARG, "bad rollout batch");
  cudaSetDevice(h->device);
  size_t sbytes = (size_t)n * h->state_bytes;
  size_t off_ret = (sbytes + 255) & ~(size_t)255;
  int rc = ensure_bytes(h, &h->d_tmp, &h->tmp_cap, off_ret + (size_t)n * sizeof(double));
  if (rc) return rc;
  rc = ensure_bytes(h, &h->d_tmp2, &h->tmp2_cap, (size_t)n * sizeof(int));
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_tmp, states, sbytes, cudaMemcpyHostToDevice, h->stream));
  double* d_ret = (double*)((char*)h->d_tmp + off_ret);
  PhaseTimer& t = h->timers[POMCP_PHASE_ROLLOUT];
  t.reset();
  CK(cudaEventRecord(t.next(), h->stream));
  DISPATCH_PID(h->prob.problem_id, {
    k_rollout_batch<PID><<<(int)((n + 255) / 256), 256, 0, h->stream>>>(
        h->dm, (const typename ModelT<PID>::State*)h->d_tmp, n, steps, (uint32_t)seed, (uint32_t)(seed >> 32), d_ret,
        (int*)h->d_tmp2);
  });
  rc = launch_check(h, "k_rollout_batch");
  if (rc) return rc;
  CK(cudaEventRecord(t.next(), h->stream));
  t.launches = 1;
  CK(cudaMemcpyAsync(ret, d_ret, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (steps_out) CK(cudaMemcpyAsync(steps_out, h->d_tmp2, (size_t)n * sizeof(int), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return POMCP_OK;
}

int32_t pomcp_pf_weights(pomcp_handle* h, const void* states, int64_t n, int32_t a, uint64_t obs_bits, uint64_t seed,
                         double* w_out, void* sp_out) {
  if (!h || !states || !w_out || n <= 0 || a < 0 || a >= h->nA) return fail(h, POMCP_ERR_INVALID_ARG, "bad pf_weights arguments");
  cudaSetDevice(h->device);
  size_t sbytes = (size_t)n * h->state_bytes;
  size_t off2 = (sbytes + 255) & ~(size_t)255;
  int rc = ensure_bytes(h, &h->d_tmp, &h->tmp_cap, off2 + sbytes);
  if (rc) return rc;
  rc = ensure_bytes(h, &h->d_tmp2, &h->tmp2_cap, (size_t)n * sizeof(double));
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_tmp, states, sbytes, cudaMemcpyHostToDevice, h->stream));
  void* d_sp = (char*)h->d_tmp + off2;
  DISPATCH_PID(h->prob.problem_id, {
    auto src = make_src<PID>(h, h->d_tmp, a, obs_bits, seed);
    k_pf_weights<PID><<<(int)((n + 255) / 256), 256, 0, h->stream>>>(src, n, (double*)h->d_tmp2,
                                                                    (typename ModelT<PID>::State*)d_sp);
  });
  rc = launch_check(h, "k_pf_weights");
  if (rc) return rc;
  CK(cudaMemcpyAsync(w_out, h->d_tmp2, (size_t)n * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  if (sp_out) CK(cudaMemcpyAsync(sp_out, d_sp, sbytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return POMCP_OK;
}

// generate_sor / initial state for the outer simulation loop (the L4 simulator of SURVEY.md §3.5 runs the
// real environment with the same model code as the planner).
int32_t pomcp_generate_sor(pomcp_handle* h, const void* state, int32_t a, uint64_t seed, uint32_t step,
                           void* next_state, uint64_t* obs_bits, double* reward, int32_t* terminal) {
  if (!h || !state || !next_state || a < 0 || a >= h->nA) return fail(h, POMCP_ERR_INVALID_ARG, "bad generate_sor arguments");
  cudaSetDevice(h->device);
  int rc = ensure_bytes(h, &h->d_tmp, &h->tmp_cap, 512);
  if (rc) return rc;
  CK(cudaMemcpyAsync(h->d_tmp, state, (size_t)h->state_bytes, cudaMemcpyHostToDevice, h->stream));
  char host[64];
  DISPATCH_PID(h->prob.problem_id, {
    EnvStep<PID>* d_out = (EnvStep<PID>*)((char*)h->d_tmp + 256);
    k_generate_sor<PID><<<1, 32, 0, h->stream>>>(h->dm, (const typename ModelT<PID>::State*)h->d_tmp, a, (uint32_t)seed,
                                                 (uint32_t)(seed >> 32), step, d_out);
    rc = launch_check(h, "k_generate_sor");
    if (rc) return rc;
    static_assert(sizeof(EnvStep<PID>) <= sizeof(host), "EnvStep too large");
    CK(cudaMemcpyAsync(host, d_out, sizeof(EnvStep<PID>), cudaMemcpyDeviceToHost, h->stream));
    CK(cudaStreamSynchronize(h->stream));
    const EnvStep<PID>* e = (const EnvStep<PID>*)host;
    std::memcpy(next_state, &e->sp, (size_t)h->state_bytes);
    if (obs_bits) *obs_bits = e->obs;
    if (reward) *reward = e->r;
    if (terminal) *terminal = e->terminal;
  });
  return POMCP_OK;
}

int32_t pomcp_root_node(pomcp_handle* h, int64_t* node_id) {
  if (!h || !node_id) return POMCP_ERR_INVALID_ARG;
  *node_id = (int64_t)h->root;
  return POMCP_OK;
}

int32_t pomcp_initial_state(pomcp_handle* h, uint64_t seed, void* state_out) {
  if (!h || !state_out) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  int rc = ensure_bytes(h, &h->d_tmp, &h->tmp_cap, 512);
  if (rc) return rc;
  DISPATCH_PID(h->prob.problem_id, {
    k_initial_state<PID><<<1, 32, 0, h->stream>>>(h->dm, (uint32_t)seed, (uint32_t)(seed >> 32),
                                                  (typename ModelT<PID>::State*)h->d_tmp);
  });
  rc = launch_check(h, "k_initial_state");
  if (rc) return rc;
  CK(cudaMemcpyAsync(state_out, h->d_tmp, (size_t)h->state_bytes, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  return POMCP_OK;
}

int32_t pomcp_dump_tree(pomcp_handle* h, int64_t cap, int64_t* n_nodes, int64_t* node_N, int64_t* act_N, double* act_V,
                        int32_t* child) {
  if (!h || !n_nodes) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  CK(cudaMemcpy(h->h_ctr, h->pool.ctr, sizeof(DevCtr), cudaMemcpyDeviceToHost));
  long long nn = (long long)std::min<unsigned long long>(h->h_ctr->node_count, h->pool.max_nodes);
  *n_nodes = nn;
  long long n = std::min<long long>(nn, cap);
  if (n <= 0) return POMCP_OK;
  const int nA = h->nA, nO = h->nO;
  std::vector<uint32_t> t32((size_t)n * nA * std::max(nO, 1));
  std::vector<uint32_t> m32((size_t)n * nA);
  std::vector<double> tv((size_t)n * nA);
  if (node_N) {
    CK(cudaMemcpy(t32.data(), h->pool.node_N, (size_t)n * 4, cudaMemcpyDeviceToHost));
    for (long long i = 0; i < n; ++i) node_N[i] = t32[(size_t)i];
  }
  if (act_N) {
    CK(cudaMemcpy(t32.data(), h->pool.act_N, (size_t)n * nA * 4, cudaMemcpyDeviceToHost));
    for (long long i = 0; i < n * nA; ++i) act_N[i] = t32[(size_t)i];
  }
  if (act_V) {
    CK(cudaMemcpy(tv.data(), h->pool.act_V, (size_t)n * nA * 8, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(m32.data(), h->pool.act_M, (size_t)n * nA * 4, cudaMemcpyDeviceToHost));
    for (long long i = 0; i < n * nA; ++i) {
      double v = tv[(size_t)i];
      if (!h->exact) v = (m32[(size_t)i] > 0 && std::isfinite(h->sp.init_V)) ? v / (double)m32[(size_t)i] : h->sp.init_V;
      act_V[i] = v;
    }
  }
  if (child && nO > 0) {
    CK(cudaMemcpy(t32.data(), h->pool.child, (size_t)n * nA * nO * 4, cudaMemcpyDeviceToHost));
    for (long long i = 0; i < n * nA * nO; ++i) child[i] = t32[(size_t)i] ? (int32_t)(t32[(size_t)i] & CHILD_ID_MASK) : -1;
  }
  return POMCP_OK;
}

int32_t pomcp_get_counters(pomcp_handle* h, pomcp_counters* out) {
  if (!h || !out) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  CK(cudaMemcpyAsync(h->h_ctr, h->pool.ctr, sizeof(DevCtr), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  const DevCtr& c = *h->h_ctr;
  out->simulations = (int64_t)c.sims;
  out->terminal_skips = (int64_t)c.skips;
  out->rollouts = (int64_t)c.rollouts;
  out->rollout_steps = (int64_t)c.rollout_steps;
  out->tree_steps = (int64_t)c.tree_steps;
  out->nodes = (int64_t)c.node_count;
  out->log_entries = (int64_t)c.log_count;
  out->max_depth_seen = (int64_t)c.max_depth;
  out->waves = h->last_waves;
  out->kernel_launches = h->launches;
  return POMCP_OK;
}

// Diagnostics: SM cycles summed over the batched workers since the counters were reset:
// [0] descent, [1] child insertion, [2] leaf rollouts, [3] backup + particle pushes,
// [4] waiting for admission, [5] admitted life time (includes fetching query ids).
int32_t pomcp_worker_cycles(pomcp_handle* h, uint64_t* out8) {
  if (!h || !out8) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  CK(cudaMemcpyAsync(h->h_ctr, h->pool.ctr, sizeof(DevCtr), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int i = 0; i < 8; ++i) out8[i] = h->h_ctr->dbg[i];
  return POMCP_OK;
}

int32_t pomcp_reset_counters(pomcp_handle* h) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  size_t off = offsetof(DevCtr, sims), end = offsetof(DevCtr, scan_ticket);
  CK(cudaMemsetAsync((char*)h->pool.ctr + off, 0, end - off, h->stream));
  CK(cudaMemsetAsync((char*)h->pool.ctr + offsetof(DevCtr, dbg), 0, sizeof(((DevCtr*)0)->dbg), h->stream));
  CK(cudaStreamSynchronize(h->stream));
  h->launches = 0;
  return POMCP_OK;
}

int32_t pomcp_last_phase_ms(pomcp_handle* h, int32_t phase, double* ms, int64_t* launches) {
  if (!h || phase < 0 || phase >= POMCP_PHASE_COUNT) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  CK(cudaStreamSynchronize(h->stream));
  if (ms) *ms = h->timers[phase].total_ms();
  if (launches) *launches = h->timers[phase].launches;
  return POMCP_OK;
}

int32_t pomcp_event_record(pomcp_handle* h, int32_t slot) {
  if (!h || slot < 0 || slot >= 8) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (!h->user_ev[slot]) CK(cudaEventCreate(&h->user_ev[slot]));
  CK(cudaEventRecord(h->user_ev[slot], h->stream));
  return POMCP_OK;
}

int32_t pomcp_event_elapsed_ms(pomcp_handle* h, int32_t s0, int32_t s1, double* ms_out) {
  if (!h || s0 < 0 || s0 >= 8 || s1 < 0 || s1 >= 8 || !ms_out || !h->user_ev[s0] || !h->user_ev[s1])
    return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  CK(cudaEventSynchronize(h->user_ev[s1]));
  float ms = 0;
  CK(cudaEventElapsedTime(&ms, h->user_ev[s0], h->user_ev[s1]));
  *ms_out = ms;
  return POMCP_OK;
}

int32_t pomcp_flush_l2(pomcp_handle* h) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (!h->d_flush) {
    h->flush_n = (size_t)256 * 1024 * 1024 / sizeof(float4);  // 256 MiB > 126 MB of L2
    CK(cudaMalloc((void**)&h->d_flush, h->flush_n * sizeof(float4)));
  }
  k_flush_l2<<<148 * 8, 256, 0, h->stream>>>(h->d_flush, h->flush_n);
  return launch_check(h, "k_flush_l2");
}

// ---- device-resident batched episodes (episode.cuh)
__global__ void k_ep_sir_done_one(TreeCtl* tc, EpisodeCtl* ec, const PfCtl* ctl, const void* new_bel, long long n) {
  if (threadIdx.x != 0 || blockIdx.x != 0 || !ec->update_pending) return;
  const int st = (int)ctl->status;
  if (st != POMCP_OK) { ec->status = st; ec->active = 0; return; }
  tc->bel = new_bel;
  tc->n_bel = n;
}

int32_t pomcp_run_episodes(pomcp_handle* h, const pomcp_episode_params* ep, int32_t n_episodes, double* rewards_out,
                           int32_t* steps_out, int32_t* status_out) {
  if (!h || !ep || !rewards_out) return fail(h, POMCP_ERR_INVALID_ARG, "null argument");
  cudaSetDevice(h->device);
  if (h->exact) return fail(h, POMCP_ERR_UNSUPPORTED, "episodes run on the batched workers (search_mode = batched)");
  if (n_episodes < 1 || n_episodes > h->n_trees)
    return fail(h, POMCP_ERR_INVALID_ARG, "n_episodes must be in 1..n_trees (one tree per episode; create the planner with n_trees >= n_episodes)");
  if (ep->max_steps < 1 || !(ep->sim_eps >= 0.0)) return fail(h, POMCP_ERR_INVALID_ARG, "bad max_steps / sim_eps");
  if (h->sp.dpw_action || h->sp.dpw_observation)
    return fail(h, POMCP_ERR_UNSUPPORTED, "device episodes with progressive widening are not supported");
  const int T = n_episodes;
  const int mode = h->exact2 ? POMCP_BELIEF_EXACT2 : (h->unweighted ? POMCP_BELIEF_UNWEIGHTED : POMCP_BELIEF_EXTERNAL);
  const long long Q = h->sp.tree_queries;
  const long long sir_n = mode == POMCP_BELIEF_EXTERNAL ? ep->sir_particles : 0;
  if (mode == POMCP_BELIEF_EXTERNAL && (sir_n < 1 || sir_n > (1ll << 20)))
    return fail(h, POMCP_ERR_INVALID_ARG, "sir_particles must be in 1..2^20 per episode");
  if (mode == POMCP_BELIEF_EXACT2 && (!(ep->initial_p1 >= 0.0) || !(ep->initial_p1 <= 1.0)))
    return fail(h, POMCP_ERR_INVALID_ARG, "initial_p1 must be a probability");
  int rc;
  if (h->pending) { rc = finish_search(h); (void)rc; }
  if (!h->d_eps) {
    CK(cudaMalloc((void**)&h->d_eps, sizeof(EpisodeCtl) * (size_t)h->n_trees));
    CK(cudaMalloc((void**)&h->d_ep_logctr, sizeof(unsigned long long) * (size_t)h->n_trees));
  }
  // belief buffers: SIR ping-pong sets of sir_n particles per episode; unweighted: room for the particles of a root
  long long bel_cap = 0;
  if (mode == POMCP_BELIEF_EXTERNAL) bel_cap = sir_n;
  if (mode == POMCP_BELIEF_UNWEIGHTED) bel_cap = std::min<long long>(Q * (long long)ep->max_steps, (long long)(h->pool.log_cap / (unsigned long long)T));
  for (int k = 0; k < (mode == POMCP_BELIEF_EXTERNAL ? 2 : (mode == POMCP_BELIEF_UNWEIGHTED ? 1 : 0)); ++k) {
    rc = ensure_bytes(h, &h->d_ep_bel[k], &h->ep_bel_cap[k], (size_t)bel_cap * (size_t)T * (size_t)h->state_bytes);
    if (rc) return rc;
  }
  h->bel_kind = 1;
  rc = fresh_root(h);  // wipes what earlier trees used; roots = nodes 0 .. n_trees - 1
  if (rc) return rc;
  EpisodeCfg ec{};
  ec.n = T; ec.max_steps = ep->max_steps; ec.initial_state = ep->initial_state; ec.updater = mode;
  ec.sim_eps = ep->sim_eps; ec.initial_p1 = ep->initial_p1;
  ec.env_k0 = (uint32_t)ep->env_seed; ec.env_k1 = (uint32_t)(ep->env_seed >> 32);
  ec.key0 = (uint32_t)h->sp.seed; ec.key1 = (uint32_t)(h->sp.seed >> 32);
  ec.rank0 = (uint32_t)(h->sp.rank * h->n_trees);
  ec.sir_n = sir_n;
  ec.log_seg = h->pool.log_cap / (unsigned long long)T;
  const int eb = (T + 127) / 128;
  DISPATCH_PID(h->prob.problem_id, {
    k_ep_init<PID><<<eb, 128, 0, h->stream>>>(h->pool, h->dm, ec, h->d_trees, h->d_eps, h->d_ep_logctr, h->d_ep_bel[0]);
    if (mode == POMCP_BELIEF_EXTERNAL)
      k_ep_init_particles<PID><<<(int)(((long long)T * sir_n + 255) / 256), 256, 0, h->stream>>>(
          h->dm, ec, (typename ModelT<PID>::State*)h->d_ep_bel[0]);
  });
  rc = launch_check(h, "k_ep_init");
  if (rc) return rc;
  PhaseTimer& ts = h->timers[POMCP_PHASE_SEARCH];
  ts.reset();
  CK(cudaEventRecord(ts.next(), h->stream));
  int cur = 0;
  for (int step = 1; step <= ep->max_steps; ++step) {
    k_ep_prepare<<<eb, 128, 0, h->stream>>>(h->d_trees, h->d_eps, T, h->epoch + (uint32_t)step);
    rc = reset_search_flags(h);
    if (rc) return rc;
    SearchCfg cfg = make_cfg(h);
    cfg.n_trees = T;
    cfg.bel = nullptr; cfg.n_bel = 0;
    cfg.exact2 = mode == POMCP_BELIEF_EXACT2 ? 1 : 0;
    rc = launch_batched_workers(h, cfg, Q);
    if (rc) return rc;
    k_root_result<<<T, 32, 0, h->stream>>>(h->pool, h->nA, h->d_trees, 0, h->sp.init_V, h->d_res);
    DISPATCH_PID(h->prob.problem_id, {
      k_ep_step<PID><<<eb, 128, 0, h->stream>>>(h->pool, h->dm, ec, h->d_trees, h->d_eps, h->d_res, (uint32_t)step);
    });
    rc = launch_check(h, "k_ep_step");
    if (rc) return rc;
    ts.launches += 4;
    if (mode == POMCP_BELIEF_UNWEIGHTED) {
      if (h->state_bytes == 4) k_ep_gather<uint32_t><<<T, 256, 0, h->stream>>>(h->pool, h->d_trees, h->d_eps, (uint32_t*)h->d_ep_bel[0], bel_cap);
      else k_ep_gather<LDState><<<T, 256, 0, h->stream>>>(h->pool, h->d_trees, h->d_eps, (LDState*)h->d_ep_bel[0], bel_cap);
      rc = launch_check(h, "k_ep_gather");
      if (rc) return rc;
    } else if (mode == POMCP_BELIEF_EXTERNAL) {
      // update(::SIRParticleFilter, b, a, o) per episode, action and observation read on the device
      for (int e = 0; e < T; ++e) {
        const uint64_t seed = ep->env_seed * 0x9E3779B97F4A7C15ull + ((uint64_t)e << 20) + (uint64_t)step;
        const Philox4 pr = philox4x32_10(0u, TAG_RESAMPLE, 0u, 0u, (uint32_t)seed, (uint32_t)(seed >> 32));
        DISPATCH_PID(h->prob.problem_id, {
          using State = typename ModelT<PID>::State;
          EpisodeSrc<PID> src;
          src.states = (const State*)h->d_ep_bel[cur] + (size_t)e * (size_t)sir_n;
          src.m = h->dm; src.ep = h->d_eps + e; src.k0 = (uint32_t)seed; src.k1 = (uint32_t)(seed >> 32);
          State* dst = (State*)h->d_ep_bel[1 - cur] + (size_t)e * (size_t)sir_n;
          rc = run_resample(h, src, sir_n, pr.w[0], sir_n, dst, false);
          if (rc == POMCP_OK) k_ep_sir_done_one<<<1, 32, 0, h->stream>>>(h->d_trees + e, h->d_eps + e, h->d_pfctl, dst, sir_n);
        });
        if (rc) return rc;
      }
      cur ^= 1;
      // every tree is discarded after its decision (src/solver.jl:278-281): recycle the pool
      k_ep_pool_wipe<<<148 * 8, 256, 0, h->stream>>>(h->pool, h->nA, h->nO, (uint32_t)h->sp.init_N, h->sp.init_V);
      k_ep_pool_restart<<<eb, 128, 0, h->stream>>>(h->pool, h->d_trees, T);
      rc = launch_check(h, "k_ep_pool_wipe");
      if (rc) return rc;
    }
  }
  CK(cudaEventRecord(ts.next(), h->stream));
  std::vector<EpisodeCtl> host((size_t)T);
  CK(cudaMemcpyAsync(host.data(), h->d_eps, sizeof(EpisodeCtl) * (size_t)T, cudaMemcpyDeviceToHost, h->stream));
  CK(cudaMemcpyAsync(h->h_ctr, h->pool.ctr, sizeof(DevCtr), cudaMemcpyDeviceToHost, h->stream));
  CK(cudaStreamSynchronize(h->stream));
  for (int e = 0; e < T; ++e) {
    rewards_out[e] = host[(size_t)e].total;
    if (steps_out) steps_out[e] = host[(size_t)e].steps;
    if (status_out) status_out[e] = host[(size_t)e].status;
  }
  h->epoch += (uint32_t)ep->max_steps + 1u;
  h->nodes_used = std::max<unsigned long long>(h->h_ctr->node_count, (unsigned long long)h->n_trees);
  h->log_used = 0;
  h->bel_kind = 0;  // the trees and beliefs of the episodes are gone: set a belief before the next search
  if (h->h_ctr->pool_overflow == 2) return fail(h, POMCP_ERR_CUDA, "worker watchdog fired");
  return POMCP_OK;
}

#ifdef POMCP_PROBE
// developer probe build only (scripts/build_variant.sh probe -DPOMCP_PROBE): per-CTA timestamps of the last one-barrier SIR launch
int32_t pomcp_debug_pf_probe(pomcp_handle* h, unsigned long long* out, int32_t n_words) {
  if (!h || !out) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  CK(cudaStreamSynchronize(h->stream));
  CK(cudaMemcpyFromSymbol(out, g_pf_probe, sizeof(unsigned long long) * (size_t)std::min(n_words, PF_MAX_GRID * 4)));
  return POMCP_OK;
}
#endif

// ---- K10: root-parallel merge over NCCL
int32_t pomcp_nccl_unique_id(void* id_out) {
  std::string err;
  if (!id_out) return POMCP_ERR_INVALID_ARG;
  if (!g_nccl.load(err)) return fail(nullptr, POMCP_ERR_NCCL, err);
  nccl_uid id;
  int rc = g_nccl.GetUniqueId(&id);
  if (rc != 0) return fail(nullptr, POMCP_ERR_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "ncclGetUniqueId");
  std::memcpy(id_out, &id, sizeof(id));
  return POMCP_OK;
}

int32_t pomcp_comm_init(pomcp_handle* h, const void* uid, int32_t rank, int32_t nranks) {
  if (!h || !uid || rank < 0 || rank >= nranks) return fail(h, POMCP_ERR_INVALID_ARG, "bad comm arguments");
  cudaSetDevice(h->device);
  if (!g_nccl.load(h->err)) return POMCP_ERR_NCCL;
  nccl_uid id;
  std::memcpy(&id, uid, sizeof(id));
  int rc = g_nccl.CommInitRank(&h->comm, nranks, id, rank);
  if (rc != 0) return fail(h, POMCP_ERR_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(rc) : "ncclCommInitRank");
  if (((long long)rank + 1) * h->n_trees > 256) return fail(h, POMCP_ERR_INVALID_ARG, "(rank + 1) * n_trees must be <= 256");
  h->nranks = nranks;
  h->sp.rank = rank;  // distinct Philox streams per tree
  return POMCP_OK;
}

int32_t pomcp_search_rootparallel(pomcp_handle* h, int64_t Q, int32_t* action_out, int64_t* N, double* V, int32_t nA) {
  if (!h) return POMCP_ERR_INVALID_ARG;
  cudaSetDevice(h->device);
  if (!h->comm) return fail(h, POMCP_ERR_INVALID_ARG, "pomcp_comm_init has not been called");
  int rc = enqueue_search(h, Q, nullptr, 0, nullptr, 0);
  if (rc) return rc;
  // one fused all-reduce of [N_a ..., N_a*V_a ..., exists_a ...] (3|A| doubles) before the arg-max of src/solver.jl:60-70
  // the cross-rank merge gets its own slot: the trees' own results stay readable (pomcp_get_tree_result)
  SearchResult* dec = h->d_res + h->n_trees + 1;
  CK(cudaMemcpyAsync(dec, h->d_res + h->res_index, sizeof(SearchResult), cudaMemcpyDeviceToDevice, h->stream));
  h->res_index = h->n_trees + 1;
  PhaseTimer& tm = h->timers[POMCP_PHASE_MERGE];
  tm.reset();
  CK(cudaEventRecord(tm.next(), h->stream));
  int nrc = g_nccl.AllReduce(dec->merge, dec->merge, (size_t)(3 * h->nA), NCCL_FLOAT64, NCCL_SUM, h->comm, h->stream);
  if (nrc != 0) return fail(h, POMCP_ERR_NCCL, g_nccl.GetErrorString ? g_nccl.GetErrorString(nrc) : "ncclAllReduce");
  k_root_merge_finish<<<1, 32, 0, h->stream>>>(h->nA, h->sp.init_V, dec);
  rc = launch_check(h, "k_root_merge_finish");
  if (rc) return rc;
  CK(cudaEventRecord(tm.next(), h->stream));
  tm.launches = 2;
  rc = finish_search(h);
  if (rc == POMCP_OK) copy_out(h, action_out, N, V, nA);
  return rc;
}

}  // extern "C"

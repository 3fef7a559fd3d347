// K9 (second half) — pool compaction after re-rooting.  The reference keeps the sub-tree under the
// new root alive and lets Julia's GC drop the rest (src/updater.jl:20-26); here the node pool and the
// particle log are append-only, so a long episode needs the dead part reclaimed:
//   1. mark + rank: a node is kept iff walking its parent links reaches the new root; the ranks of the
//      kept nodes (allocation order) come out of the same single-pass scan that evaluates the marks, and
//      the new root then swaps ranks with whichever kept node got rank 0;
//   2. copy the kept nodes' statistics, with child ids remapped, into a scratch block and back to the
//      front of the pool; re-initialise the freed tail; rebuild the observation hash table (LightDark);
//   3. filter the particle log the same way (stable, node ids remapped).
#pragma once
#include <cstdint>
#include "pf.cuh"
#include "tree.cuh"

namespace pomcp {

constexpr uint32_t NO_NODE = 0xffffffffu;

struct SubtreeSrc {
  const uint32_t* node_pslot;
  uint32_t new_root;
  uint32_t nA;
  __device__ __forceinline__ unsigned long long operator()(long long i) const {
    uint32_t n = (uint32_t)i;
    // Walk the parent links until the new root or the pool's first node (id 0: its own parent, like every
    // detached or never-used id).  Ids mostly grow along a path, so nodes older than the new root could stop
    // at once, but the batched workers insert with ids allocated ahead of time and a child can be older than
    // its parent.
    while (n != new_root && n != 0u) n = node_pslot[n] / nA;
    return n == new_root ? 1ull : 0ull;
  }
};
struct NewIdSink {
  uint32_t* newid;
  __device__ __forceinline__ void operator()(long long i, unsigned long long inc, unsigned long long v) const {
    newid[i] = v ? (uint32_t)(inc - 1) : NO_NODE;
  }
};
__global__ void __launch_bounds__(SCAN_THREADS) k_rank_subtree(const uint32_t* node_pslot, uint32_t new_root,
                                                               uint32_t nA, long long n_nodes, uint32_t* newid,
                                                               unsigned long long* status,
                                                               unsigned long long* ticket) {
  SubtreeSrc s{node_pslot, new_root, nA};
  NewIdSink k{newid};
  chained_scan(s, k, n_nodes, status, ticket);
}

// The new root must become node 0.  Ranks follow allocation order, and with the batched workers' pre-allocated
// ids a descendant can be older than the new root and take rank 0: swap the two ranks.
__global__ void k_root_rank(const uint32_t* newid, uint32_t new_root, uint32_t* root_rank) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *root_rank = newid[new_root];
}
__global__ void k_root_first(uint32_t* newid, long long n_nodes, uint32_t new_root, const uint32_t* root_rank) {
  const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  const uint32_t r = *root_rank;
  if (i >= n_nodes || r == 0u) return;
  if ((uint32_t)i == new_root) newid[i] = 0u;
  else if (newid[i] == 0u) newid[i] = r;
}

// scratch copy of the kept nodes (SoA mirrors of the pool arrays)
struct PoolScratch {
  uint32_t* node_N;
  uint32_t* node_flags;
  uint32_t* node_pslot;
  unsigned long long* node_obs;
  uint32_t* act_N;
  uint32_t* act_M;
  double* act_V;
  uint32_t* child;
  uint32_t* node_amask;  // action widening only
};

__global__ void k_copy_kept(const __grid_constant__ Pool P, const __grid_constant__ PoolScratch T,
                            const uint32_t* newid, uint32_t new_root, int nA, int nO, long long n_nodes) {
  long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (n >= n_nodes) return;
  uint32_t m = newid[n];
  if (m == NO_NODE) return;
  T.node_N[m] = P.node_N[n];
  T.node_flags[m] = P.node_flags[n];
  uint32_t ps = P.node_pslot[n];
  // parent slot -> (new parent id, same action); the new root has no parent
  T.node_pslot[m] = ((uint32_t)n == new_root) ? 0u : newid[ps / (uint32_t)nA] * (uint32_t)nA + ps % (uint32_t)nA;
  if (P.node_amask) T.node_amask[m] = P.node_amask[n];
  if (P.node_obs) T.node_obs[m] = P.node_obs[n];
  else if (P.node_bel) T.node_obs[m] = (unsigned long long)__double_as_longlong(P.node_bel[n]);  // same scratch column
  for (int a = 0; a < nA; ++a) {
    size_t so = (size_t)n * nA + a, sn = (size_t)m * nA + a;
    T.act_N[sn] = P.act_N[so];
    T.act_M[sn] = P.act_M[so];
    T.act_V[sn] = P.act_V[so];
    for (int o = 0; o < nO; ++o) {
      uint32_t c = P.child[so * nO + o];
      T.child[sn * nO + o] = c ? (newid[c & CHILD_ID_MASK] | (c & CHILD_EXPANDED)) : 0u;
    }
  }
}

// continuous observations: re-insert the kept nodes into the emptied (slot, obs) -> id table
__global__ void k_rehash(const __grid_constant__ Pool P, long long n_kept) {
  long long n = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (n <= 0 || n >= n_kept) return;  // node 0 is the root: it is nobody's child
  uint64_t h = mix64((uint64_t)P.node_pslot[n] * 0x9E3779B97F4A7C15ull ^ P.node_obs[n]) & P.ht_mask;
  for (;;) {
    uint32_t old = atomicCAS(&P.ht[h], 0u, (uint32_t)n);
    if (old == 0) return;
    h = (h + 1) & P.ht_mask;
  }
}

struct LogKeepSrc {
  const uint32_t* log_node;
  const uint32_t* newid;
  __device__ __forceinline__ unsigned long long operator()(long long i) const {
    return newid[log_node[i]] != NO_NODE ? 1ull : 0ull;
  }
};
template <class State>
struct LogCompactSink {
  const uint32_t* log_node;
  const State* log_state;
  const uint32_t* newid;
  uint32_t* out_node;
  State* out_state;
  __device__ __forceinline__ void operator()(long long i, unsigned long long inc, unsigned long long v) const {
    if (v) {
      out_node[inc - 1] = newid[log_node[i]];
      out_state[inc - 1] = log_state[i];
    }
  }
};
template <class State>
__global__ void __launch_bounds__(SCAN_THREADS) k_log_compact(const uint32_t* log_node, const State* log_state,
                                                              const uint32_t* newid, long long n, uint32_t* out_node,
                                                              State* out_state, unsigned long long* status,
                                                              unsigned long long* ticket) {
  LogKeepSrc s{log_node, newid};
  LogCompactSink<State> k{log_node, log_state, newid, out_node, out_state};
  chained_scan(s, k, n, status, ticket);
}

// fill [first, first+count) of the pool with fresh ActNode(a, init_N, init_V) values
__global__ void k_pool_fill_range(const __grid_constant__ Pool P, int nA, int nO, unsigned long long first,
                                  unsigned long long count, uint32_t init_N, double init_V, int exact) {
  unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x;
  unsigned long long stride = (unsigned long long)gridDim.x * blockDim.x;
  double v0 = exact ? init_V : (init_N > 0 && fabs(init_V) != INFINITY ? init_V * (double)init_N : 0.0);
  for (unsigned long long k = i; k < count * nA; k += stride) {
    P.act_N[first * nA + k] = init_N;
    P.act_M[first * nA + k] = init_N;
    P.act_V[first * nA + k] = v0;
  }
  for (unsigned long long k = i; k < count; k += stride) {
    P.node_N[first + k] = 0;
    P.node_flags[first + k] = 0;
    P.node_pslot[first + k] = 0;
    if (P.node_obs) P.node_obs[first + k] = 0;
    if (P.node_amask) P.node_amask[first + k] = 0;
  }
  if (P.child)
    for (unsigned long long k = i; k < count * nA * nO; k += stride) P.child[first * nA * nO + k] = 0;
}

__global__ void k_set_counts(DevCtr* c, unsigned long long node_total, unsigned long long log_total) {
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    c->node_count = node_total;
    c->log_count = log_total;
  }
}

}  // namespace pomcp

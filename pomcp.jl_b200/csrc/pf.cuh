// K6/K7/K8 — SIR particle-filter update on the device, and K9's particle gather.
// Replaces ParticleFilters.jl's SimpleParticleFilter.update + LowVarianceResampler.resample
// (not vendored in the reference; call site test/runtests.jl:57, semantics SURVEY.md App. B.3)
// and the per-node ParticleCollection of the unweighted filter (src/particle_filter.jl:27-30).
//
// Two implementations of the same integers (so the same indices, bit for bit):
//  * k_pf_fused: one persistent cooperative launch for sets that fit one co-resident grid, weights
//    kept in registers across two grid barriers (latency-bound sizes);
//  * streaming path for any size, three passes shaped for HBM and for the issue slots:
//      K6 k_pf_weigh      read s_i                      -> write w_i (fp64, -1 = terminal), max weight, first alive
//         k_pf_total      read w_i                      -> T = sum of q_i = trunc(w_i * 2^s)   (s needs the max)
//      K7+K8 k_pf_scan_emit  read w_i, single-pass decoupled-look-back scan of q_i with 4 consecutive
//                         particles per thread, closed-form output range of every particle from
//                         (C_i, T), read s_i of the particles that are copied, write the copies
//    (3S + 24 bytes per particle against the algorithmic 2S + 16: the state is read twice because the
//    output ranges need T, which needs the scale, which needs the largest weight.)
// Fixed point makes the cumulative sums associative, so the systematic-resampling indices are
// identical for every scan order (the oracle walks the same integers sequentially).
#pragma once
#include <cstdint>
#include "models.cuh"
#include "tree.cuh"

namespace pomcp {

constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 4;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;
constexpr unsigned long long SCAN_VAL_MASK = (1ull << 62) - 1;

struct PfCtl {
  unsigned long long max_bits;     // bits of the largest weight (weights >= 0: order-preserving)
  unsigned long long first_alive_inv;  // ~(index of the first non-terminal particle); 0 = none yet
  unsigned long long total;        // T = sum q_i at the scale in use
  unsigned int any_alive, status;
  unsigned long long total_ceil;   // streaming path: sum q_i at the a-priori scale (pf_scale_ceil)
  unsigned int over_ceil, _pad;    // a weight above the model's analytic ceiling was seen: the a-priori scale is off
};
// Internal status of the one-barrier kernels: the a-priori scale does not resolve these weights (or a weight is
// above the ceiling) - the caller runs the max-scaled kernels instead.  Never crosses the ABI.
constexpr int PF_RETRY_MAX_SCALE = 100;
// Scale rule (the oracle's resample_fixed states the same): weights of a model are bounded by an analytic
// ceiling 2^(E+1) (probabilities: E = 0; LightDark's Normal pdf with sigma >= 0.01: 39.9 < 2^6, E = 5), so
// s_ceil = 61 - ceil(log2 n) - E keeps sum q < 2^62 without looking at the data, and the max-weight pass with its
// grid barrier goes away.  It is used when no weight exceeds the ceiling and the total keeps at least 40 bits
// (sum q >= 2^40); otherwise the scale comes from the largest weight as before.
constexpr unsigned long long PF_MIN_TOTAL_CEIL = 1ull << 40;
constexpr int PF_NO_CEIL = -100000;

__device__ __forceinline__ unsigned long long ld_volatile_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(v) : "l"(p));
  return v;
}

// Decoupled look-back (Merrill & Garland): warp 0 of the tile publishes its aggregate, then
// polls 32 predecessors at a time until it meets an inclusive prefix.
__device__ __forceinline__ unsigned long long scan_lookback(unsigned long long* status, int tile,
                                                            unsigned long long aggregate) {
  const int lane = threadIdx.x & 31;
  if (tile == 0) {
    if (lane == 0) atomicExch(&status[0], (2ull << 62) | aggregate);
    return 0ull;
  }
  if (lane == 0) atomicExch(&status[tile], (1ull << 62) | aggregate);
  unsigned long long exclusive = 0;
  int pos = tile - 1;
  for (;;) {
    int idx = pos - lane;
    unsigned long long st = (idx >= 0) ? ld_volatile_u64(&status[idx]) : (2ull << 62);
    unsigned flag = (unsigned)(st >> 62);
    unsigned pref = __ballot_sync(0xffffffffu, flag == 2u);
    unsigned none = __ballot_sync(0xffffffffu, flag == 0u);
    int p = pref ? (__ffs(pref) - 1) : -1;
    unsigned need = (p >= 0) ? ((p == 31) ? 0xffffffffu : ((2u << p) - 1u)) : 0xffffffffu;
    if (none & need) continue;  // a needed predecessor has not published yet
    unsigned long long v = ((need >> lane) & 1u) ? (st & SCAN_VAL_MASK) : 0ull;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    exclusive += v;
    if (p >= 0) break;
    pos -= 32;
  }
  if (lane == 0) atomicExch(&status[tile], (2ull << 62) | ((exclusive + aggregate) & SCAN_VAL_MASK));
  return exclusive;
}

// Generic single-pass inclusive scan: src(i) -> u64 value (< 2^62 in total), sink(i, inclusive, value).
// Items are warp-striped so that every load/store instruction is coalesced.
template <class Src, class Sink>
__device__ __forceinline__ void chained_scan(Src& src, Sink& sink, long long n, unsigned long long* status,
                                             unsigned long long* ticket) {
  __shared__ unsigned s_tile;
  __shared__ unsigned long long s_warp[SCAN_THREADS / 32];
  __shared__ unsigned long long s_excl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = (unsigned)atomicAdd(ticket, 1ull);
  __syncthreads();
  const int tile = (int)s_tile;
  const long long base = (long long)tile * SCAN_TILE + (long long)warp * 32 * SCAN_ITEMS;
  unsigned long long v[SCAN_ITEMS], inc[SCAN_ITEMS];
  unsigned long long carry = 0;
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    long long i = base + j * 32 + lane;
    v[j] = (i < n) ? src(i) : 0ull;
    unsigned long long x = v[j];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += t;
    }
    inc[j] = x + carry;
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) s_warp[warp] = carry;
  __syncthreads();
  if (warp == 0) {
    unsigned long long w = (lane < SCAN_THREADS / 32) ? s_warp[lane] : 0ull;
    unsigned long long x = w;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += t;
    }
    unsigned long long aggregate = __shfl_sync(0xffffffffu, x, 31);
    if (lane < SCAN_THREADS / 32) s_warp[lane] = x - w;  // exclusive over warps
    unsigned long long excl = scan_lookback(status, tile, aggregate);
    if (lane == 0) s_excl = excl;
  }
  __syncthreads();
  const unsigned long long off0 = s_excl + s_warp[warp];
#pragma unroll
  for (int j = 0; j < SCAN_ITEMS; ++j) {
    long long i = base + j * 32 + lane;
    if (i < n) sink(i, off0 + inc[j], v[j]);
  }
}

// Particle sets live in cudaMalloc'ed arrays, so a 16-byte state (LightDark: {status, y}, 8-byte alignment as a C
// struct) can move as ONE 16-byte access: as two 8-byte ones every state cost two sector requests each way (ncu on
// k_pf_emit_tiled, 2^24 LightDark particles: 2.0 sectors per 16-byte copy, L2 throughput 73 %).
template <class S>
__device__ __forceinline__ S pf_ld_state(const S* p) {
  if constexpr (sizeof(S) == 16) {
    const uint4 v = *reinterpret_cast<const uint4*>(p);
    S s;
    memcpy(&s, &v, 16);
    return s;
  } else {
    return *p;
  }
}
template <class S>
__device__ __forceinline__ void pf_st_state(S* p, const S& s) {
  if constexpr (sizeof(S) == 16) {
    uint4 v;
    memcpy(&v, &s, 16);
    *reinterpret_cast<uint4*>(p) = v;
  } else {
    *p = s;
  }
}

// ---- weight sources
template <int PID>
struct ModelSrc {
  using M = ModelT<PID>;
  using State = typename M::State;
  const State* states;
  DevModel m;
  int a;
  uint64_t obs;
  uint32_t k0, k1;
  static constexpr int CEIL_EXP = (PID == POMCP_LIGHTDARK1D) ? 5 : 0;  // weights < 2^(CEIL_EXP + 1), see pf_scale rule
  // Weight classes (tiled streaming path): for a fixed (a, o) the weight of a discrete problem depends on a few bits of
  // the propagated state only - the state itself (Tiger, Baby), the cell and the sensed rock's bit (RockSample) - so the
  // QUANTISED weights live in a small shared-memory table built per launch and a particle's q is one look-up: nothing
  // is stored per particle between the two passes.  0 = no such table (continuous observations).
  static constexpr int NCLASS = (PID == POMCP_TIGER || PID == POMCP_BABY) ? 2 : (PID == POMCP_ROCKSAMPLE ? 512 : 0);
  // The class can be read off the INPUT state: RockSample's sensing actions leave the state alone and every other
  // action's weight is the same for all states; Tiger's listening likewise, and an opened door weighs 0.5 whatever the
  // state.  (Baby's next state is drawn: its class needs the propagated particle.)
  static constexpr bool CLASS_IN = PID == POMCP_ROCKSAMPLE || PID == POMCP_TIGER;
  // the action leaves every (non-terminal) state as it is: RockSample's sensing actions, Tiger's listening
  __device__ __forceinline__ bool identity() const {
    return (PID == POMCP_ROCKSAMPLE && a >= 5) || (PID == POMCP_TIGER && a == 0);
  }
  __device__ __forceinline__ int wclass(const State& sp) const {
    if constexpr (PID == POMCP_ROCKSAMPLE) {
      const int ri = a >= 5 ? a - 5 : 0;
      return (int)((sp & 0xffu) | (((sp >> (8 + ri)) & 1u) << 8));
    } else if constexpr (PID == POMCP_TIGER || PID == POMCP_BABY) {
      return (int)(sp & 1u);
    } else {
      return 0;
    }
  }
  __device__ __forceinline__ double class_weight(int c) const {  // obs_weight of any state of class c
    if constexpr (PID == POMCP_ROCKSAMPLE) {
      const int ri = a >= 5 ? a - 5 : 0;
      const uint32_t sp = (uint32_t)(c & 0xff) | ((uint32_t)((c >> 8) & 1) << (8 + ri));
      return M::obs_weight(m, a, sp, obs);
    } else if constexpr (PID == POMCP_TIGER || PID == POMCP_BABY) {
      return M::obs_weight(m, a, (uint32_t)c, obs);
    } else {
      return 0.0;
    }
  }
  __device__ __forceinline__ bool skip() const { return false; }
  // propagate + weight one particle: generate_s then pdf(observation(a, sp), o)
  __device__ __forceinline__ bool eval(long long i, State& sp, double& w) const {
    State s = pf_ld_state(states + i);
    if (M::terminal(m, s)) { sp = s; w = 0.0; return false; }  // dropped before weighting
    double ut0 = 0.0;
    if (M::T_DRAW) ut0 = u01(philox4x32_10(0u, TAG_PF, (uint32_t)i, (uint32_t)((unsigned long long)i >> 32), k0, k1).w[0]);
    uint64_t o;
    double r;
    M::template step<false, double>(m, s, a, ut0, 0.0, 0.0, sp, o, r);
    w = M::obs_weight(m, a, sp, obs);
    return true;
  }
  // the same from a state that is already in registers
  __device__ __forceinline__ bool eval_from(long long i, const State& s, State& sp, double& w) const {
    if (M::terminal(m, s)) { sp = s; w = 0.0; return false; }
    double ut0 = 0.0;
    if (M::T_DRAW) ut0 = u01(philox4x32_10(0u, TAG_PF, (uint32_t)i, (uint32_t)((unsigned long long)i >> 32), k0, k1).w[0]);
    uint64_t o;
    double r;
    M::template step<false, double>(m, s, a, ut0, 0.0, 0.0, sp, o, r);
    w = M::obs_weight(m, a, sp, obs);
    return true;
  }
  // generate_s only (the emit phase needs the state, not the weight, of the particles it copies)
  __device__ __forceinline__ void propagate(long long i, State& sp) const { propagate_from(i, pf_ld_state(states + i), sp); }
  __device__ __forceinline__ State load(long long i) const { return pf_ld_state(states + i); }
  __device__ __forceinline__ void propagate_from(long long i, const State& s, State& sp) const {
    double ut0 = 0.0;
    if (M::T_DRAW) ut0 = u01(philox4x32_10(0u, TAG_PF, (uint32_t)i, (uint32_t)((unsigned long long)i >> 32), k0, k1).w[0]);
    uint64_t o;
    double r;
    M::template step<false, double>(m, s, a, ut0, 0.0, 0.0, sp, o, r);
  }
};

struct ArraySrc {  // replayed weights (parity hook): payload is the index itself
  using State = long long;
  static constexpr int CEIL_EXP = PF_NO_CEIL;  // arbitrary weights: the scale always comes from the largest one
  static constexpr int NCLASS = 0;
  static constexpr bool CLASS_IN = false;
  __device__ __forceinline__ bool identity() const { return false; }
  __device__ __forceinline__ int wclass(const State&) const { return 0; }
  __device__ __forceinline__ double class_weight(int) const { return 0.0; }
  __device__ __forceinline__ bool skip() const { return false; }
  const double* w_in;
  __device__ __forceinline__ void propagate(long long i, State& sp) const { sp = i; }
  __device__ __forceinline__ State load(long long i) const { return i; }
  __device__ __forceinline__ void propagate_from(long long, const State& s, State& sp) const { sp = s; }
  __device__ __forceinline__ bool eval(long long i, State& sp, double& w) const {
    sp = i;
    w = w_in[i];
    return true;
  }  __device__ __forceinline__ bool eval_from(long long i, const State&, State& sp, double& w) const { return eval(i, sp, w); }
};

__device__ __forceinline__ int pf_scale(unsigned long long max_bits, long long n) {
  int e = (int)((max_bits >> 52) & 0x7ffull) - 1023;
  int L = (n <= 1) ? 0 : 64 - __clzll(n - 1);  // ceil(log2 n)
  return 61 - L - e;
}
__device__ __forceinline__ int pf_scale_ceil(int ceil_exp, long long n) {
  int L = (n <= 1) ? 0 : 64 - __clzll(n - 1);
  return 61 - L - ceil_exp;
}
__device__ __forceinline__ bool pf_over_ceil(double w, int ceil_exp) {  // w >= 2^(ceil_exp + 1), inf or NaN
  return !(w < u2d((unsigned long long)(1023 + ceil_exp + 1) << 52));
}
// trunc(w * 2^s) like ldexp + conversion, without the library call: scaling by a power of two is exact
// (s in [-994, 1135]; above 1000 in two exact steps, since 2^s itself would overflow), and wherever
// a product rounds (denormal results) it is below 1 and truncates to 0 either way.
__device__ __forceinline__ double pf_pow2(int k) { return u2d((unsigned long long)(1023 + k) << 52); }
__device__ __forceinline__ unsigned long long pf_quantise(double w, int s) {
  const int s1 = s < 1000 ? s : 1000;
  double x = w * pf_pow2(s1);
  if (s > 1000) x *= pf_pow2(s - 1000);
  return (w > 0.0) ? __double2ull_rz(x) : 0ull;
}

// K6 (streaming): propagate + weight every particle once, keep the weight.  Striped so that every
// load and store is a full-line coalesced access; terminal particles are stored as -1.
template <class Src>
__global__ void __launch_bounds__(256) k_pf_weigh(const __grid_constant__ Src src, long long n, double* w_out,
                                                  PfCtl* ctl) {
  __shared__ unsigned long long s_max[8];
  __shared__ unsigned long long s_first[8];
  // n < 2^32 (checked at the ABI): 32-bit indices; the first alive particle of a thread is the first it meets
  unsigned long long mx = 0, first = ~0ull;
  const uint32_t stride = gridDim.x * blockDim.x, nn = (uint32_t)n;
  uint32_t first32 = 0xffffffffu;
#pragma unroll 4
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < nn; i += stride) {
    typename Src::State sp;
    double wk;
    const bool alive = src.eval((long long)i, sp, wk);
    wk = alive ? (wk > 0.0 ? wk : 0.0) : -1.0;
    w_out[i] = wk;
    if (alive) {
      first32 = min(first32, i);
      const unsigned long long bts = d2u(wk);
      mx = bts > mx ? bts : mx;
    }
    if (i + stride < i) break;  // index wrap guard for n close to 2^32
  }
  if (first32 != 0xffffffffu) first = first32;
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    unsigned long long om = __shfl_xor_sync(0xffffffffu, mx, off), of = __shfl_xor_sync(0xffffffffu, first, off);
    mx = om > mx ? om : mx;
    first = of < first ? of : first;
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_max[warp] = mx; s_first[warp] = first; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) {
      mx = s_max[k] > mx ? s_max[k] : mx;
      first = s_first[k] < first ? s_first[k] : first;
    }
    if (first != ~0ull) {
      atomicMax(&ctl->max_bits, mx);
      atomicMax(&ctl->first_alive_inv, ~first);
      ctl->any_alive = 1;
    }
  }
}

// T = sum_i trunc(w_i * 2^s): the output ranges of K8 need the total before the scan reaches the end.
// With a weight ceiling (ceil_exp != PF_NO_CEIL) the total at the a-priori scale is summed in the same pass; the
// scan picks one of the two (pf_pick_scale).
__device__ __forceinline__ bool pf_use_ceil(const PfCtl* ctl, int ceil_exp) {
  if (ceil_exp == PF_NO_CEIL) return false;
  const int e = (int)((ctl->max_bits >> 52) & 0x7ffull) - 1023;
  return e <= ceil_exp && ctl->total_ceil >= PF_MIN_TOTAL_CEIL;
}
__global__ void __launch_bounds__(256) k_pf_total(const double* w, long long n, PfCtl* ctl, int ceil_exp) {
  __shared__ unsigned long long s_sum[8], s_sum2[8];
  const int scale = pf_scale(ctl->max_bits, n);
  const bool both = ceil_exp != PF_NO_CEIL && ((int)((ctl->max_bits >> 52) & 0x7ffull) - 1023) <= ceil_exp;
  const int scale_c = pf_scale_ceil(both ? ceil_exp : 0, n);
  unsigned long long sum = 0, sum_c = 0;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += stride) {
    const double wi = w[i];
    sum += pf_quantise(wi, scale);
    if (both) sum_c += pf_quantise(wi, scale_c);
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    sum += __shfl_xor_sync(0xffffffffu, sum, off);
    sum_c += __shfl_xor_sync(0xffffffffu, sum_c, off);
  }
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) { s_sum[warp] = sum; s_sum2[warp] = sum_c; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < (int)(blockDim.x >> 5); ++k) { sum += s_sum[k]; sum_c += s_sum2[k]; }
    if (sum) atomicAdd(&ctl->total, sum);
    if (sum_c) atomicAdd(&ctl->total_ceil, sum_c);
  }
}

// Number of systematic points U_m = (ru/2^32 + m) * T / n_out, m in [0, n_out), with U_m <= C, i.e.
//   M(C) = floor((C*n_out*2^32 - ru*T) / (2^32*T)) + 1   (0 when the numerator is negative).
// Write C*n_out = k*T + rem (0 <= rem < T).  Then M(C) = k + [rem*2^32 >= ru*T] = k + [rem >= thr]
// with thr = ceil(ru*T / 2^32).  k < 2^32 comes from one fp64 multiply (relative error 2^-51, so at
// most one off) and is corrected with rem computed in wrapping 64-bit arithmetic: the true
// remainder of a k that is one off lies in (-T, 2T), well inside int64 since T < 2^62.
struct PfDiv {
  unsigned long long T, thr;
  long long n_out;
  double ratio;  // n_out / T
  __device__ __forceinline__ PfDiv(unsigned long long T_, uint32_t ru, long long n_out_) : T(T_), n_out(n_out_) {
    thr = (unsigned long long)(((unsigned __int128)ru * T_ + 0xffffffffull) >> 32);
    ratio = (double)n_out_ / (double)T_;
  }
  __device__ __forceinline__ long long points_below(unsigned long long C) const {
    unsigned long long k = (unsigned long long)((double)C * ratio);
    long long rem = (long long)(C * (unsigned long long)n_out - k * T);
    while (rem < 0) { --k; rem += (long long)T; }
    while ((unsigned long long)rem >= T) { ++k; rem -= (long long)T; }
    long long cnt = (long long)k + ((unsigned long long)rem >= thr ? 1 : 0);
    return cnt < n_out ? cnt : n_out;
  }
  // The same count from one fp64 evaluation, used when it is provably the exact integer: x = C * n_out / T - ru / 2^32
  // is computed with a relative error below 2^-50 (u64 -> fp64 conversion, the rounded ratio, one product, one
  // difference), so floor(x) + 1 is M(C) unless x lies within eps = (|x| + 1) * 2^-46 of an integer; those few
  // (about 2^-20 of the particles at n_out = 2^20) take the integer path.  Same integers either way.
  __device__ __forceinline__ long long points_below_fast(unsigned long long C, double ru_frac) const {
    const double x = (double)C * ratio - ru_frac;
    const double fl = floor(x);
    const double fr = x - fl;
    const double eps = (fabs(x) + 1.0) * 1.4210854715202004e-14;  // 2^-46
    if (fr > eps && fr < 1.0 - eps && x < 4.0e18) {
      const long long cnt = x < 0.0 ? 0ll : (long long)fl + 1ll;
      return cnt < n_out ? cnt : n_out;
    }
    return points_below(C);
  }
  // The same test with ONE tolerance for the whole launch: |x| <= n_out + 1, so eps = (n_out + 2) * 2^-46 covers every
  // particle, and "x is further than eps from an integer" is floor(x - eps) == floor(x + eps) - two conversions with
  // round-down instead of floor, a difference, a product and three comparisons.  Counts fit 32 bits (n_out < 2^32).
  struct Fast {
    double sub_lo, sub_hi;  // ru / 2^32 + eps, ru / 2^32 - eps
  };
  __device__ __forceinline__ Fast fast(uint32_t ru) const {
    const double f = (double)ru * 2.3283064365386963e-10, eps = ((double)n_out + 2.0) * 1.4210854715202004e-14;
    return Fast{f + eps, f - eps};
  }
  __device__ __forceinline__ long long points_below_fast2(unsigned long long C, const Fast& f) const {
    const double y = (double)C * ratio;  // <= n_out < 2^32 (plus rounding)
    const double a = y - f.sub_lo, b = y - f.sub_hi;  // x - eps, x + eps
    if (a >= 0.0 && b < 4294967295.0) {
      const uint32_t fa = __double2uint_rd(a), fb = __double2uint_rd(b);
      if (fa == fb) {
        const long long cnt = (long long)fa + 1ll;
        return cnt < n_out ? cnt : n_out;
      }
    } else if (b < 0.0) {
      return 0ll;
    }
    return points_below(C);
  }
  // Branch-free form for a batch of independent evaluations: returns the count (32 bits: n_out < 2^32) and clears `ok`
  // where the fp64 evaluation does not prove it (the caller then asks points_below for that C).
  __device__ __forceinline__ uint32_t points_below_try(unsigned long long C, const Fast& f, uint32_t n_out32, bool& ok) const {
    const double y = (double)C * ratio;
    const double a = y - f.sub_lo, b = y - f.sub_hi;  // x - eps, x + eps
    const uint32_t fa = __double2uint_rd(a), fb = __double2uint_rd(b);  // saturating conversions
    const bool neg = b < 0.0;
    ok = ok && (neg || (a >= 0.0 && b < 4294967295.0 && fa == fb));
    const uint32_t cnt = neg ? 0u : min(fa, n_out32 - 1u) + 1u;  // n_out >= 1
    return cnt;
  }
  // The same count for a run of cumulative sums C, C + q1, C + q1 + q2, ...: (k, rem) is carried and a
  // step adds q * n_out to the remainder, so the quotient grows by the number of copies of that
  // particle (a few subtractions) instead of being divided out again.  Falls back to the closed form
  // when q * n_out does not fit or the particle takes many copies.
  struct Run {
    unsigned long long C, k, rem;
  };
  __device__ __forceinline__ Run start(unsigned long long C) const {
    unsigned long long k = (unsigned long long)((double)C * ratio);
    long long rem = (long long)(C * (unsigned long long)n_out - k * T);
    while (rem < 0) { --k; rem += (long long)T; }
    while ((unsigned long long)rem >= T) { ++k; rem -= (long long)T; }
    return Run{C, k, (unsigned long long)rem};
  }
  __device__ __forceinline__ long long count(const Run& r) const {
    long long cnt = (long long)r.k + (r.rem >= thr ? 1 : 0);
    return cnt < n_out ? cnt : n_out;
  }
  __device__ __forceinline__ void advance(Run& r, unsigned long long q) const {
    r.C += q;
    const unsigned long long d = q * (unsigned long long)n_out;
    if (__umul64hi(q, (unsigned long long)n_out) == 0ull && d < (3ull << 62)) {
      unsigned long long x = r.rem + d;  // rem < T < 2^62 and d < 3 * 2^62: no wrap
      unsigned long long k = r.k;
#pragma unroll
      for (int it = 0; it < 3; ++it)
        if (x >= T) { x -= T; ++k; }
      if (x < T) { r.rem = x; r.k = k; return; }
    }
    r = start(r.C);
  }
};

// K7+K8 (streaming): single-pass scan of the quantised weights fused with the resampling.  A thread owns
// PF2_ITEMS consecutive particles (16-byte loads, a thread-local prefix and one u64 warp scan per
// thread instead of one per particle); particle i owns the output range [M(C_{i-1}), M(C_i)) and
// writes that many copies of its propagated state.  Ranges of neighbouring particles are adjacent,
// so the writes of a warp land in a few consecutive lines.
constexpr int PF2_THREADS = 256;
constexpr int PF2_ITEMS = 8;
constexpr int PF2_TILE = PF2_THREADS * PF2_ITEMS;

template <class Src, class Out>
__global__ void __launch_bounds__(PF2_THREADS) k_pf_scan_emit(const __grid_constant__ Src src, long long n,
                                                              const double* __restrict__ w, PfCtl* ctl, uint32_t ru,
                                                              long long n_out, Out* out, unsigned long long* status,
                                                              unsigned long long* ticket) {
  __shared__ unsigned s_tile;
  __shared__ unsigned long long s_warp[PF2_THREADS / 32];
  __shared__ unsigned long long s_excl;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_tile = (unsigned)atomicAdd(ticket, 1ull);
  __syncthreads();
  const int tile = (int)s_tile;
  const unsigned long long max_bits = ctl->max_bits;
  const bool use_ceil = pf_use_ceil(ctl, Src::CEIL_EXP);
  const unsigned long long T = use_ceil ? ctl->total_ceil : ctl->total;
  if (tile == 0 && threadIdx.x == 0) {
    int st = POMCP_OK;
    if (!ctl->any_alive) st = POMCP_ERR_ALL_SAMPLES_TERMINAL;  // "all states in the particle collection were terminal"
    else if (max_bits == 0 || T == 0) st = POMCP_ERR_PARTICLE_DEPLETION;
    ctl->status = st;
  }
  if (T == 0) return;  // uniform: nothing to resample
  const int scale = use_ceil ? pf_scale_ceil(Src::CEIL_EXP, n) : pf_scale(max_bits, n);
  const long long base = (long long)tile * PF2_TILE + (long long)threadIdx.x * PF2_ITEMS;
  double wv[PF2_ITEMS];
  if (base + PF2_ITEMS <= n) {
    const double2* p = reinterpret_cast<const double2*>(w + base);  // base is a multiple of PF2_ITEMS: 64-byte aligned
#pragma unroll
    for (int k = 0; k < PF2_ITEMS / 2; ++k) {
      const double2 a = p[k];
      wv[2 * k] = a.x;
      wv[2 * k + 1] = a.y;
    }
  } else {
#pragma unroll
    for (int k = 0; k < PF2_ITEMS; ++k) wv[k] = (base + k < n) ? w[base + k] : -1.0;
  }
  // small states are requested now and used after the scan (16-byte states would cost 32 registers)
  constexpr bool PRELOAD = sizeof(typename Src::State) <= 8;
  typename Src::State st_in[PRELOAD ? PF2_ITEMS : 1];
  if (PRELOAD) {
#pragma unroll
    for (int k = 0; k < PF2_ITEMS; ++k)
      if (base + k < n) st_in[PRELOAD ? k : 0] = src.load(base + k);
  }
  unsigned long long inc[PF2_ITEMS];
  unsigned long long run = 0;
#pragma unroll
  for (int k = 0; k < PF2_ITEMS; ++k) {
    run += pf_quantise(wv[k], scale);
    inc[k] = run;
  }
  // exclusive prefix of the thread totals inside the warp
  unsigned long long x = run;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
    if (lane >= off) x += t;
  }
  if (lane == 31) s_warp[warp] = x;
  const unsigned long long thread_excl = x - run;
  __syncthreads();
  if (warp == 0) {
    unsigned long long v = (lane < PF2_THREADS / 32) ? s_warp[lane] : 0ull, y = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, y, off);
      if (lane >= off) y += t;
    }
    const unsigned long long aggregate = __shfl_sync(0xffffffffu, y, 31);
    if (lane < PF2_THREADS / 32) s_warp[lane] = y - v;  // exclusive over warps
    const unsigned long long excl = scan_lookback(status, tile, aggregate);
    if (lane == 0) s_excl = excl;
  }
  __syncthreads();
  const unsigned long long off0 = s_excl + s_warp[warp] + thread_excl;  // C just before this thread's particles

  const unsigned long long first_alive = ~ctl->first_alive_inv;
  const PfDiv dv(T, ru, n_out);
  PfDiv::Run rn = dv.start(off0);
  long long lo = dv.count(rn);
  unsigned long long prev_inc = 0;
#pragma unroll
  for (int k = 0; k < PF2_ITEMS; ++k) {
    const long long i = base + k;
    dv.advance(rn, inc[k] - prev_inc);
    prev_inc = inc[k];
    const long long hi = dv.count(rn);
    if (i < n && wv[k] >= 0.0) {
      if ((unsigned long long)i == first_alive) lo = 0;
      const uint32_t copies = (uint32_t)(hi - lo);  // n_out < 2^32
      if (copies) {
        typename Src::State sp;
        if (PRELOAD) src.propagate_from(i, st_in[PRELOAD ? k : 0], sp);
        else src.propagate(i, sp);
        Out* o = out + lo;
        pf_st_state(o, (Out)sp);
        if (copies > 1) {
          pf_st_state(o + 1, (Out)sp);
          for (uint32_t j = 2; j < copies; ++j) pf_st_state(o + j, (Out)sp);
        }
      }
    }
    lo = hi;
  }
}

// ---- Tiled streaming path (round 2): the a-priori scale (pf_scale_ceil) needs no look at the data, so the weights
// are quantised where they are computed and the scan needs no look-back:
//   K6  k_pf_weigh_q      read s_i, propagate + weight, q_i = trunc(w_i * 2^s_ceil) -> write q_i, one sum per tile of
//                         2048 particles (flags: alive particles, first alive particle, a weight above the ceiling)
//       k_pf_tile_prefix  one CTA: exclusive prefix of the tile sums, T, the status word (or PF_RETRY_MAX_SCALE when
//                         the a-priori scale does not resolve these weights: the caller then runs the max-scaled path)
//   K7+K8 k_pf_emit_tiled read q_i (and s_i), scan inside the tile, output range [M(C_{i-1}), M(C_i)) of every
//                         particle from the fp64 form of M (exact, integer fall-back), write the copies
// 3S + 16 bytes per particle (max-scaled path: 3S + 24), every tile independent of every other.  Same integers as
// every other path, so the same indices.
// Sources with a weight-class table (the discrete problems) run it as TWO launches: nothing is stored per particle (the
// emit kernel looks the class weight up again: 3S bytes per particle), a CTA owns a contiguous chunk of tiles in both
// kernels, k_pf_weigh_q publishes one sum per CTA and k_pf_emit_tiled derives total, status and offsets from those.
// Quantised weights of the weight classes of `src` at `scale` (see ModelSrc::NCLASS); returns whether one of them
// is above the ceiling.  Ends with __syncthreads().
template <class Src>
__device__ __forceinline__ bool pf_fill_qtab(const Src& src, int scale, unsigned long long* qtab) {
  bool over = false;
  for (int c = threadIdx.x; c < Src::NCLASS; c += blockDim.x) {
    const double w = src.class_weight(c);
    over = over || pf_over_ceil(w, Src::CEIL_EXP);
    qtab[c] = pf_quantise(w, scale);
  }
  __syncthreads();
  return over;
}

template <class Src>
__global__ void __launch_bounds__(PF2_THREADS) k_pf_weigh_q(const __grid_constant__ Src src, long long n, long long ntiles,
                                                            unsigned long long* __restrict__ q_out,
                                                            unsigned long long* __restrict__ tile_sum, PfCtl* ctl) {
  __shared__ unsigned long long s_sum[PF2_THREADS / 32];
  __shared__ unsigned s_first[PF2_THREADS / 32];
  __shared__ unsigned long long qtab[Src::NCLASS > 0 ? Src::NCLASS : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int scale = pf_scale_ceil(Src::CEIL_EXP, n);
  bool over = false;
  if constexpr (Src::NCLASS > 0) over = pf_fill_qtab(src, scale, qtab);  // once per CTA, which then walks several tiles
  unsigned long long first_cta = ~0ull;  // thread 0: first alive particle of the tiles of this CTA
  // Sources with a weight-class table: a CTA owns a CONTIGUOUS chunk of tiles and publishes ONE sum, so k_pf_emit_tiled
  // (same grid, same chunks) gets its offset from the <= 8 * SMs chunk sums and carries it from tile to tile - there is
  // no prefix kernel between the two passes (one CTA, 25 us for the 32 K tiles of 2^26 particles).
  constexpr bool CHUNK = Src::NCLASS > 0;
  const long long per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
  const long long t0 = CHUNK ? (long long)blockIdx.x * per_cta : (long long)blockIdx.x;
  const long long t1 = CHUNK ? (t0 + per_cta < ntiles ? t0 + per_cta : ntiles) : ntiles;
  const long long tstep = CHUNK ? 1 : (long long)gridDim.x;
  unsigned long long acc_sum = 0, acc_first = ~0ull;  // CHUNK: this thread's sum / first alive particle over the chunk
  for (long long tile = t0; tile < t1; tile += tstep) {
    const long long base = tile * PF2_TILE;
    const bool full = base + PF2_TILE <= n;  // uniform
    unsigned long long sum = 0;
    unsigned first = 0xffffffffu;  // offset inside the tile of the first alive particle this thread met
    constexpr bool VEC = Src::NCLASS > 0 && sizeof(typename Src::State) == 4 && PF2_ITEMS == 8;
    if constexpr (VEC) {
      // Full tile of 4-byte states: both 16-byte loads of a thread are requested before anything looks at a state, and
      // the look-ups are selects, not branches.  (The general loop below compiles to load -> branch on the terminal bit
      // -> look-up -> next load: eight memory round trips in a row per thread; 2^26 RockSample particles took 138 us.)
      if (full) {
        const uint4* p4 = reinterpret_cast<const uint4*>(src.states + base);  // tiles are 8 KB: 16-byte aligned
        const uint4 v0 = p4[threadIdx.x], v1 = p4[PF2_THREADS + threadIdx.x];
        const uint32_t w8[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
#pragma unroll
        for (int k = 0; k < 8; ++k) {
          typename Src::State s0;
          memcpy(&s0, &w8[k], 4);
          const unsigned off = (unsigned)(((k >> 2) * PF2_THREADS + (int)threadIdx.x) * 4 + (k & 3));
          const bool alive = !Src::M::terminal(src.m, s0);
          unsigned long long q;
          if constexpr (Src::CLASS_IN) {
            q = qtab[src.wclass(s0)];  // the class index is in range for any bit pattern
          } else {
            typename Src::State sp;
            src.propagate_from(base + off, s0, sp);
            q = qtab[src.wclass(sp)];
          }
          sum += alive ? q : 0ull;
          first = alive ? min(first, off) : first;
        }
      }
    }
    // striped over the CTA: every load and store instruction of a warp covers consecutive particles
#pragma unroll
    for (int k = 0; k < PF2_ITEMS; ++k) {
      if (VEC && full) break;  // uniform
      const long long i = base + (long long)k * PF2_THREADS + threadIdx.x;
      if (full || i < n) {
        typename Src::State sp;
        unsigned long long q = 0;
        bool alive;
        if constexpr (Src::NCLASS > 0) {
          const typename Src::State s0 = src.load(i);
          alive = !Src::M::terminal(src.m, s0);
          if constexpr (Src::CLASS_IN) {
            q = alive ? qtab[src.wclass(s0)] : 0ull;
          } else if (alive) {
            src.propagate_from(i, s0, sp);
            q = qtab[src.wclass(sp)];
          }
        } else {
          double wk;
          alive = src.eval(i, sp, wk);
          if (alive) {
            over = over || pf_over_ceil(wk, Src::CEIL_EXP);
            q = pf_quantise(wk, scale);
          }
          q_out[i] = q;
        }
        if (alive) first = min(first, (unsigned)(k * PF2_THREADS + (int)threadIdx.x));
        sum += q;
      }
    }
    if constexpr (CHUNK) {
      // one sum per CTA: a thread keeps its own sum (and first alive particle) over the whole chunk and the CTA reduces
      // once, after the loop - no shuffles and no barrier per tile, the warps stream their tiles independently
      acc_sum += sum;
      if (first != 0xffffffffu && acc_first == ~0ull) acc_first = (unsigned long long)(base + first);  // tiles ascend
      continue;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      sum += __shfl_xor_sync(0xffffffffu, sum, off);
      first = min(first, __shfl_xor_sync(0xffffffffu, first, off));
    }
    if (lane == 0) {
      s_sum[warp] = sum;
      s_first[warp] = first;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < PF2_THREADS / 32; ++k) {
        sum += s_sum[k];
        first = min(first, s_first[k]);
      }
      tile_sum[tile] = sum;
      if (first != 0xffffffffu && first_cta == ~0ull) first_cta = (unsigned long long)(base + first);  // tiles ascend
    }
    __syncthreads();
  }
  if constexpr (CHUNK) {
    __shared__ unsigned long long s_first64[PF2_THREADS / 32];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      acc_sum += __shfl_xor_sync(0xffffffffu, acc_sum, off);
      const unsigned long long of = __shfl_xor_sync(0xffffffffu, acc_first, off);
      acc_first = of < acc_first ? of : acc_first;
    }
    if (lane == 0) {
      s_sum[warp] = acc_sum;
      s_first64[warp] = acc_first;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      for (int k = 1; k < PF2_THREADS / 32; ++k) {
        acc_sum += s_sum[k];
        acc_first = s_first64[k] < acc_first ? s_first64[k] : acc_first;
      }
      tile_sum[blockIdx.x] = acc_sum;  // (also the CTAs beyond the last tile: zero)
      first_cta = acc_first;
    }
  }
  over = __any_sync(0xffffffffu, over);
  if (lane == 0 && over) ctl->over_ceil = 1;
  if (threadIdx.x == 0 && first_cta != ~0ull) {
    atomicMax(&ctl->first_alive_inv, ~first_cta);
    ctl->any_alive = 1;
  }
}

// Exclusive prefix of the tile sums in place, grand total and status.  One CTA of 32 warps; per round a LANE owns 16
// consecutive tile sums (eight 16-byte loads, its own running sum, eight 16-byte stores) and a warp 512: one u64
// warp scan per 512 tiles and one CTA scan per 16 K.  (A thread per contiguous chunk of the whole array took 59 us for
// 32 K tiles - dependent loads; a warp scan per row of 32 took 28 us - 32 scans per warp; 32 sums per lane need 64
// registers for the values alone, which 1024 threads do not have: the kernel spilled and took 28 us as well.)
constexpr int PFX_PER = 16;
__global__ void __launch_bounds__(1024) k_pf_tile_prefix(unsigned long long* tile_sum, long long ntiles, PfCtl* ctl) {
  __shared__ unsigned long long s_seg[32];
  __shared__ unsigned long long s_carry;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (long long round0 = 0; round0 < ntiles; round0 += 1024 * PFX_PER) {
    const long long first = round0 + ((long long)warp * 32 + lane) * PFX_PER;  // a multiple of 16 (tile_sum itself is 16-byte aligned)
    unsigned long long v[PFX_PER];
    if (first + PFX_PER <= ntiles) {
      const ulonglong2* p = reinterpret_cast<const ulonglong2*>(tile_sum + first);
#pragma unroll
      for (int j = 0; j < PFX_PER / 2; ++j) {
        const ulonglong2 a = p[j];
        v[2 * j] = a.x;
        v[2 * j + 1] = a.y;
      }
    } else {
#pragma unroll
      for (int j = 0; j < PFX_PER; ++j) v[j] = first + j < ntiles ? tile_sum[first + j] : 0ull;
    }
    unsigned long long mine = 0;
#pragma unroll
    for (int j = 0; j < PFX_PER; ++j) mine += v[j];
    unsigned long long x = mine;  // inclusive over the lanes of the warp
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += t;
    }
    if (lane == 31) s_seg[warp] = x;
    __syncthreads();
    unsigned long long run = s_carry + (x - mine);
#pragma unroll
    for (int k = 0; k < 32; ++k) run += (k < warp) ? s_seg[k] : 0ull;
    __syncthreads();
    if (threadIdx.x == 1023) s_carry = run + mine;  // everything up to the end of this round
#pragma unroll
    for (int j = 0; j < PFX_PER; ++j) {
      const unsigned long long t = v[j];
      v[j] = run;
      run += t;
    }
    if (first + PFX_PER <= ntiles) {
      ulonglong2* p = reinterpret_cast<ulonglong2*>(tile_sum + first);
#pragma unroll
      for (int j = 0; j < PFX_PER / 2; ++j) p[j] = make_ulonglong2(v[2 * j], v[2 * j + 1]);
    } else {
#pragma unroll
      for (int j = 0; j < PFX_PER; ++j)
        if (first + j < ntiles) tile_sum[first + j] = v[j];
    }
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    const unsigned long long T = s_carry;
    ctl->total_ceil = T;
    ctl->total = T;
    int st = POMCP_OK;
    if (!ctl->any_alive) st = POMCP_ERR_ALL_SAMPLES_TERMINAL;
    else if (ctl->over_ceil || T < PF_MIN_TOTAL_CEIL) st = PF_RETRY_MAX_SCALE;
    ctl->status = (unsigned)st;
  }
}

// Resident CTAs per SM the register allocation must allow: the kernel is bound by the latencies of a few warps per
// scheduler (ncu: 80 registers -> 3 CTAs, issue-active 50 %, long-scoreboard and fixed-latency stalls on top), and left
// to itself ptxas takes 126-130 registers.  64 registers (4 CTAs) for the 4-byte states of the weight-class sources -
// 2^26 RockSample particles: 447 -> 395 us per update -, 80 (3 CTAs) for wider states.  No spills either way.
#ifndef PF_EMIT_MIN_CTAS_CLASS
#define PF_EMIT_MIN_CTAS_CLASS 4
#endif
#ifndef PF_EMIT_MIN_CTAS_WIDE
#define PF_EMIT_MIN_CTAS_WIDE 3
#endif
template <class Src, class Out>
__global__ void __launch_bounds__(PF2_THREADS, (Src::NCLASS > 0 ? PF_EMIT_MIN_CTAS_CLASS : PF_EMIT_MIN_CTAS_WIDE))
k_pf_emit_tiled(const __grid_constant__ Src src, long long n, long long ntiles,
                                                               const unsigned long long* __restrict__ q,
                                                               const unsigned long long* __restrict__ tile_off,
                                                               PfCtl* ctl, uint32_t ru, long long n_out, Out* out) {
  __shared__ unsigned long long s_warp[PF2_THREADS / 32];
  __shared__ unsigned long long qtab[Src::NCLASS > 0 ? Src::NCLASS : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr bool CLASSES = Src::NCLASS > 0;
  constexpr bool CHUNK = CLASSES;  // k_pf_weigh_q left one sum per CTA (chunk of tiles), no prefix kernel ran
  unsigned long long T, chunk_run = 0;  // chunk_run: C before the next tile of this CTA's chunk
  if constexpr (CHUNK) {
    // Every CTA derives the grand total, the status word (the same rule as k_pf_tile_prefix) and its own offset from
    // the chunk sums: <= 8 * SMs values, a few coalesced loads per thread.
    __shared__ unsigned long long s_red[2][PF2_THREADS / 32];
    unsigned long long all = 0, before = 0;
    for (unsigned c = threadIdx.x; c < gridDim.x; c += PF2_THREADS) {
      const unsigned long long v = tile_off[c];
      all += v;
      before += c < blockIdx.x ? v : 0ull;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      all += __shfl_xor_sync(0xffffffffu, all, off);
      before += __shfl_xor_sync(0xffffffffu, before, off);
    }
    if (lane == 0) {
      s_red[0][warp] = all;
      s_red[1][warp] = before;
    }
    __syncthreads();
    all = 0;
    before = 0;
#pragma unroll
    for (int k = 0; k < PF2_THREADS / 32; ++k) {
      all += s_red[0][k];
      before += s_red[1][k];
    }
    T = all;
    chunk_run = before;
    int st = POMCP_OK;
    if (!ctl->any_alive) st = POMCP_ERR_ALL_SAMPLES_TERMINAL;
    else if (ctl->over_ceil || T < PF_MIN_TOTAL_CEIL) st = PF_RETRY_MAX_SCALE;
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      ctl->total_ceil = T;
      ctl->total = T;
      ctl->status = (unsigned)st;
    }
    if (st != POMCP_OK) return;  // uniform over the grid: all terminal, or the max-scaled path takes over
  } else {
    if (ctl->status != POMCP_OK) return;  // uniform: all terminal, or the max-scaled path takes over
    T = ctl->total_ceil;
  }
  constexpr bool CLASS_IN = CLASSES && Src::CLASS_IN;
  constexpr bool PRELOAD = sizeof(typename Src::State) <= 8;  // (16-byte states would cost 32 registers)
  // Wider states are staged through shared memory: the tile's states are requested by coalesced loads when the tile
  // starts (a load per copied particle inside the emit loop made every thread wait for eight round trips in a row);
  // one entry of padding per eight keeps a warp's reads (lane stride 8 particles) off each other's banks.
  constexpr bool STAGE = !PRELOAD;
  __shared__ typename Src::State s_st[STAGE ? PF2_TILE + PF2_TILE / 8 : 1];
  if constexpr (CLASSES) pf_fill_qtab(src, pf_scale_ceil(Src::CEIL_EXP, n), qtab);
  const unsigned long long first_alive = ~ctl->first_alive_inv;
  const PfDiv dv(T, ru, n_out);
  const PfDiv::Fast fs = dv.fast(ru);
  const bool same = src.identity();  // uniform: copies are the input states themselves
  const uint32_t n_out32 = (uint32_t)n_out;
  const long long per_cta = (ntiles + gridDim.x - 1) / gridDim.x;
  const long long t0 = CHUNK ? (long long)blockIdx.x * per_cta : (long long)blockIdx.x;
  const long long t1 = CHUNK ? (t0 + per_cta < ntiles ? t0 + per_cta : ntiles) : ntiles;
  const long long tstep = CHUNK ? 1 : (long long)gridDim.x;
  for (long long tile = t0; tile < t1; tile += tstep) {
    const long long base = tile * PF2_TILE + (long long)threadIdx.x * PF2_ITEMS;
    const bool full = (tile + 1) * PF2_TILE <= n;  // uniform: no bounds checks inside a full tile
    unsigned long long tile_base_sum = chunk_run;
    if constexpr (!CHUNK) tile_base_sum = tile_off[tile];  // requested early: used after the scan
    unsigned long long qv[PF2_ITEMS];
    typename Src::State st_in[PRELOAD ? PF2_ITEMS : 1];  // input states (propagated in place where the class needs it)
    if constexpr (CLASSES) {
      static_assert(PRELOAD, "weight classes are for small discrete states");
      if (sizeof(typename Src::State) == 4 && PF2_ITEMS == 8 && full) {  // two 16-byte loads (base is a multiple of 8)
        const uint4* p4 = reinterpret_cast<const uint4*>(src.states + base);
        const uint4 lo4 = p4[0], hi4 = p4[1];
        const uint32_t w8[8] = {lo4.x, lo4.y, lo4.z, lo4.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
        for (int k = 0; k < PF2_ITEMS; ++k) memcpy(&st_in[k], &w8[k], 4);
      } else {
#pragma unroll
        for (int k = 0; k < PF2_ITEMS; ++k)
          if (full || base + k < n) st_in[k] = src.load(base + k);
      }
#pragma unroll
      for (int k = 0; k < PF2_ITEMS; ++k) {
        qv[k] = 0ull;
        if (full || base + k < n) {
          if (!Src::M::terminal(src.m, st_in[k])) {
            if constexpr (CLASS_IN) {
              qv[k] = qtab[src.wclass(st_in[k])];
            } else {
              typename Src::State sp;
              src.propagate_from(base + k, st_in[k], sp);
              st_in[k] = sp;
              qv[k] = qtab[src.wclass(sp)];
            }
          }
        }
      }
    } else {
      if (base + PF2_ITEMS <= n) {
        const ulonglong2* p = reinterpret_cast<const ulonglong2*>(q + base);  // 64-byte aligned
#pragma unroll
        for (int k = 0; k < PF2_ITEMS / 2; ++k) {
          const ulonglong2 a = p[k];
          qv[2 * k] = a.x;
          qv[2 * k + 1] = a.y;
        }
      } else {
#pragma unroll
        for (int k = 0; k < PF2_ITEMS; ++k) qv[k] = (base + k < n) ? q[base + k] : 0ull;
      }
      if (PRELOAD) {  // small states are requested now and used after the scan
#pragma unroll
        for (int k = 0; k < PF2_ITEMS; ++k)
          if (base + k < n) st_in[PRELOAD ? k : 0] = src.load(base + k);
      } else {
#pragma unroll
        for (int k = 0; k < PF2_ITEMS; ++k) {
          const int idx = k * PF2_THREADS + (int)threadIdx.x;
          const long long i = tile * PF2_TILE + idx;
          if (full || i < n) s_st[STAGE ? idx + idx / 8 : 0] = src.load(i);
        }
      }
    }
    unsigned long long run = 0;
#pragma unroll
    for (int k = 0; k < PF2_ITEMS; ++k) {
      run += qv[k];
      qv[k] = run;  // inclusive inside the thread
    }
    unsigned long long x = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += t;
    }
    if (lane == 31) s_warp[warp] = x;
    __syncthreads();
    unsigned long long warp_excl = 0, tile_total = 0;
#pragma unroll
    for (int k = 0; k < PF2_THREADS / 32; ++k) {
      const unsigned long long wk = s_warp[k];
      tile_total += wk;
      warp_excl += (k < warp) ? wk : 0ull;
    }
    __syncthreads();  // s_warp is rewritten by the next tile
    const unsigned long long off0 = tile_base_sum + warp_excl + (x - run);  // C just before this thread's particles
    if constexpr (CHUNK) chunk_run += tile_total;
    // nothing of this thread's is copied: terminal or zero-weight particles only (the first alive particle owns output 0)
    if (run == 0ull && first_alive - (unsigned long long)base >= (unsigned long long)PF2_ITEMS) continue;
    // The nine range ends of this thread as one straight-line batch of independent fp64 evaluations (their conversion
    // and fp64 latencies overlap; evaluated one after another inside the branches of the copy loop they were a
    // dependent chain per particle).  A particle of zero weight has hi == lo and copies nothing; an end the fp64 form
    // does not prove (about 2^-20 of them) is recomputed in integers.
    uint32_t m[PF2_ITEMS + 1];
    bool ok = true;
    m[0] = dv.points_below_try(off0, fs, n_out32, ok);
#pragma unroll
    for (int k = 0; k < PF2_ITEMS; ++k) m[k + 1] = dv.points_below_try(off0 + qv[k], fs, n_out32, ok);
    if (!ok) {
      m[0] = (uint32_t)dv.points_below_fast2(off0, fs);
#pragma unroll
      for (int k = 0; k < PF2_ITEMS; ++k) m[k + 1] = (uint32_t)dv.points_below_fast2(off0 + qv[k], fs);
    }
    const unsigned long long fa_rel = first_alive - (unsigned long long)base;  // < PF2_ITEMS: this thread owns output 0
#pragma unroll
    for (int k = 0; k < PF2_ITEMS; ++k) {
      const long long i = base + k;
      const uint32_t lo = fa_rel == (unsigned long long)k ? 0u : m[k];
      const uint32_t copies = m[k + 1] - lo;  // n_out < 2^32
      if (copies) {
        typename Src::State sp;
        if ((CLASSES && !CLASS_IN) || (PRELOAD && same)) sp = st_in[PRELOAD ? k : 0];
        else if (PRELOAD) src.propagate_from(i, st_in[PRELOAD ? k : 0], sp);
        else src.propagate_from(i, s_st[STAGE ? (int)threadIdx.x * (PF2_ITEMS + 1) + k : 0], sp);
        // up to four copies straight-line (predicated stores): a loop here runs at a handful of active lanes
        Out* o = out + lo;
        const Out v = (Out)sp;
        pf_st_state(o, v);
        if (copies > 1) pf_st_state(o + 1, v);
        if (copies > 2) pf_st_state(o + 2, v);
        if (copies > 3) pf_st_state(o + 3, v);
        if (copies > 4)
          for (uint32_t j = 4; j < copies; ++j) pf_st_state(o + j, v);
      }
    }
  }
}

// Upstream running-sum semantics on one thread (verification mode of pomcp_resample_systematic).
__global__ void k_resample_seq_f64(const double* w, long long n, double r01, long long n_out, long long* idx,
                                   PfCtl* ctl) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double W = 0.0;
  for (long long i = 0; i < n; ++i) W += w[i];
  if (!(W > 0.0)) { ctl->status = POMCP_ERR_PARTICLE_DEPLETION; return; }
  double step = W / (double)n_out;
  double U = r01 * W / (double)n_out;
  double c = w[0];
  long long i = 0;
  for (long long mth = 0; mth < n_out; ++mth) {
    while (U > c) {
      if (i + 1 >= n) break;
      ++i;
      c += w[i];
    }
    U += step;
    idx[mth] = i;
  }
  ctl->status = POMCP_OK;
}

// ---- K6+K7+K8 fused: one persistent cooperative launch for particle sets that fit the grid
// (n <= co-resident CTAs * 4096; 148 SMs x 2 CTAs x 4096 items = 1.2 M particles on B200).
// The particles are split evenly over the grid, `items` (<= 8) consecutive particles per thread; a thread
// keeps their states and fp64 weights in registers across two grid-wide barriers, so HBM sees each input
// state once and each output state once:
//   phase A  read s_i (all of a thread's loads in flight together), propagate + weight
//                                               -> global max weight, first alive particle
//   barrier  phase B  quantise, thread-local prefix, one u64 warp scan per thread, CTA scan
//                                               -> per-CTA totals (gridDim values)
//   barrier  phase C  offsets from the CTA totals; (quotient, remainder) of C * n_out / T carried from
//                     particle to particle (PfDiv::Run), write the copies
// Same integers as the streaming path, so the same indices.
constexpr int PF_THREADS = 512;
constexpr int PF_ITEMS = 8;
constexpr int PF_CTA_ITEMS = PF_THREADS * PF_ITEMS;
constexpr int PF_MAX_GRID = 1024;                 // CTA totals that fit a control block
constexpr int PF_CTRL_WORDS = 9 + PF_MAX_GRID + 7;  // u64 words: PfCtl (8), barrier counter (1), CTA totals

// Grid barrier of a co-resident grid.  One release (fence + add) on the way in, relaxed polling - an acquire load per
// poll from ~300 CTAs on one address took the barrier 2 - 6 us after its last arrival (scripts/pf_barrier_probe.py) -
// and one fence on the way out.
__device__ __forceinline__ void pf_grid_barrier(unsigned int* bar, unsigned int target) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("red.relaxed.gpu.global.add.u32 [%0], 1;" ::"l"(bar) : "memory");
    unsigned int v;
    for (;;) {
      asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(bar) : "memory");
      if (v >= target) break;
      __nanosleep(64);
    }
    __threadfence();
  }
  __syncthreads();
}
// Sum of the CTA totals before this CTA and the grand total, by the whole CTA: every thread loads one total (two
// past 512 CTAs), so the phase costs one L2 round trip - a volatile load per loop iteration of one warp, each waited
// for, cost the emit phase ten of them (scripts/pf_barrier_probe.py).  Ends with __syncthreads(); the results are
// in s_off / s_total.
__device__ __forceinline__ void pf_cta_offsets(const unsigned long long* cta_tot, unsigned long long* s_part_before,
                                               unsigned long long* s_part_all, unsigned long long& s_off,
                                               unsigned long long& s_total) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = (int)gridDim.x, i0 = (int)threadIdx.x, i1 = (int)threadIdx.x + PF_THREADS;
  const unsigned long long t0 = i0 < g ? __ldcg(&cta_tot[i0]) : 0ull;
  const unsigned long long t1 = i1 < g ? __ldcg(&cta_tot[i1]) : 0ull;
  unsigned long long all = t0 + t1;
  unsigned long long before = (i0 < (int)blockIdx.x ? t0 : 0ull) + (i1 < (int)blockIdx.x ? t1 : 0ull);
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    before += __shfl_xor_sync(0xffffffffu, before, off);
    all += __shfl_xor_sync(0xffffffffu, all, off);
  }
  if (lane == 0) { s_part_before[warp] = before; s_part_all[warp] = all; }
  __syncthreads();
  if (warp == 0) {
    before = lane < PF_THREADS / 32 ? s_part_before[lane] : 0ull;
    all = lane < PF_THREADS / 32 ? s_part_all[lane] : 0ull;
#pragma unroll
    for (int off = 8; off > 0; off >>= 1) {
      before += __shfl_xor_sync(0xffffffffu, before, off);
      all += __shfl_xor_sync(0xffffffffu, all, off);
    }
    if (lane == 0) { s_off = before; s_total = all; }
  }
  __syncthreads();
}
static_assert(PF_MAX_GRID <= 2 * PF_THREADS, "pf_cta_offsets loads two totals per thread");
static_assert(PF_THREADS / 32 <= 16, "second-level reduction covers 16 warps");

template <class Src, class Out>
__global__ void __launch_bounds__(PF_THREADS, 2) k_pf_fused(const __grid_constant__ Src src, long long n, PfCtl* ctl,
                                                            uint32_t ru, long long n_out, Out* out,
                                                            unsigned long long* cta_tot, unsigned int* bar,
                                                            int items, unsigned long long* next_ctl) {
  using State = typename Src::State;
  __shared__ unsigned long long s_a[PF_THREADS / 32];
  __shared__ unsigned long long s_b[PF_THREADS / 32];
  __shared__ unsigned long long s_c[PF_THREADS / 32];
  __shared__ unsigned long long s_off, s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long base = ((long long)blockIdx.x * PF_THREADS + threadIdx.x) * items;
  // the other control block, for the next launch (nobody uses it now); this launch's own lines are
  // requested into L2 so that the first atomics on them do not pay a DRAM miss
  if (blockIdx.x == gridDim.x - 1)
    for (int k = threadIdx.x; k < PF_CTRL_WORDS; k += PF_THREADS) next_ctl[k] = 0ull;
  if (src.skip()) return;  // (episodes: no update pending for this belief) - the same for every CTA
  if (threadIdx.x == 0) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ctl));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(cta_tot + blockIdx.x));
  }

  // ---- phase A: propagate + weight.  States of <= 8 bytes stay in registers for phase C.
  constexpr bool KEEP = sizeof(State) <= 8;
  State st[KEEP ? PF_ITEMS : 1];
  if (KEEP) {
#pragma unroll
    for (int k = 0; k < PF_ITEMS; ++k)
      if (k < items && base + k < n) st[KEEP ? k : 0] = src.load(base + k);
  }
  double w[PF_ITEMS];
  unsigned alive_mask = 0;
  unsigned long long mx = 0, first = ~0ull;
#pragma unroll
  for (int k = 0; k < PF_ITEMS; ++k) {
    w[k] = 0.0;
    if (k < items && base + k < n) {
      State sp;
      const bool alive = KEEP ? src.eval_from(base + k, st[KEEP ? k : 0], sp, w[k]) : src.eval(base + k, sp, w[k]);
      if (alive) {
        alive_mask |= 1u << k;
        if (first == ~0ull) first = (unsigned long long)(base + k);
        const unsigned long long b = (w[k] > 0.0) ? d2u(w[k]) : 0ull;
        mx = b > mx ? b : mx;
      } else {
        w[k] = 0.0;
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    unsigned long long om = __shfl_xor_sync(0xffffffffu, mx, off), of = __shfl_xor_sync(0xffffffffu, first, off);
    mx = om > mx ? om : mx;
    first = of < first ? of : first;
  }
  if (lane == 0) { s_a[warp] = mx; s_b[warp] = first; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < PF_THREADS / 32; ++k) {
      mx = s_a[k] > mx ? s_a[k] : mx;
      first = s_b[k] < first ? s_b[k] : first;
    }
    if (first != ~0ull) {
      atomicMax(&ctl->max_bits, mx);
      atomicMax(&ctl->first_alive_inv, ~first);
      ctl->any_alive = 1;
    }
  }
  pf_grid_barrier(bar, gridDim.x);

  // ---- phase B: quantise, prefix inside the thread, scan the thread totals in the warp and the CTA
  const unsigned long long max_bits = ld_volatile_u64(&ctl->max_bits);
  const int scale = pf_scale(max_bits, n);
  unsigned long long inc[PF_ITEMS];
  unsigned long long run = 0;
#pragma unroll
  for (int k = 0; k < PF_ITEMS; ++k) {
    run += ((alive_mask >> k) & 1u) ? pf_quantise(w[k], scale) : 0ull;
    inc[k] = run;
  }
  unsigned long long x = run;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
    if (lane >= off) x += t;
  }
  const unsigned long long thread_excl = x - run;
  __syncthreads();  // s_a is re-used
  if (lane == 31) s_a[warp] = x;
  __syncthreads();
  if (warp == 0) {
    unsigned long long v = (lane < PF_THREADS / 32) ? s_a[lane] : 0ull, y = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, y, off);
      if (lane >= off) y += t;
    }
    if (lane < PF_THREADS / 32) s_a[lane] = y - v;  // exclusive over warps
    if (lane == 31) atomicExch(&cta_tot[blockIdx.x], y);
  }
  pf_grid_barrier(bar, 2u * gridDim.x);

  // ---- offsets: sum of the CTA totals before this CTA, and the grand total T
  pf_cta_offsets(cta_tot, s_b, s_c, s_off, s_total);
  const unsigned long long T = s_total;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->total = T;
    int stt = POMCP_OK;
    if (!ctl->any_alive) stt = POMCP_ERR_ALL_SAMPLES_TERMINAL;
    else if (max_bits == 0 || T == 0) stt = POMCP_ERR_PARTICLE_DEPLETION;
    ctl->status = stt;
  }
  if (T == 0) return;

  // ---- phase C: particle i owns outputs [M(C_{i-1}), M(C_i))
  const unsigned long long first_alive = ~ld_volatile_u64(&ctl->first_alive_inv);
  const unsigned long long off0 = s_off + s_a[warp] + thread_excl;  // C just before this thread's particles
  const PfDiv dv(T, ru, n_out);
  PfDiv::Run rn = dv.start(off0);
  long long lo = dv.count(rn);
  unsigned long long prev_inc = 0;
#pragma unroll
  for (int k = 0; k < PF_ITEMS; ++k) {
    const long long i = base + k;
    dv.advance(rn, inc[k] - prev_inc);
    prev_inc = inc[k];
    const long long hi = dv.count(rn);
    if ((alive_mask >> k) & 1u) {
      if ((unsigned long long)i == first_alive) lo = 0;
      const uint32_t copies = (uint32_t)(hi - lo);  // n_out < 2^32
      if (copies) {
        State sp;
        if (KEEP) src.propagate_from(i, st[KEEP ? k : 0], sp);
        else src.propagate(i, sp);
        Out* o = out + lo;
        pf_st_state(o, (Out)sp);
        if (copies > 1) {
          pf_st_state(o + 1, (Out)sp);
          for (uint32_t j = 2; j < copies; ++j) pf_st_state(o + j, (Out)sp);
        }
      }
    }
    lo = hi;
  }
}

// The same three phases with the particles striped over the warp (particle = base + row * 32 + lane; one u64
// warp scan per row, closed-form output range per particle): used for states wider than 8 bytes, whose
// loads only coalesce this way (LightDark, 16 bytes: 44 us against 51 us at 2^20 particles).
template <class Src, class Out>
__global__ void __launch_bounds__(PF_THREADS, 2) k_pf_fused_striped(const __grid_constant__ Src src, long long n, PfCtl* ctl,
                                                            uint32_t ru, long long n_out, Out* out,
                                                            unsigned long long* cta_tot, unsigned int* bar,
                                                            int rows) {
  __shared__ unsigned long long s_a[PF_THREADS / 32];
  __shared__ unsigned long long s_b[PF_THREADS / 32];
  __shared__ unsigned long long s_c[PF_THREADS / 32];
  __shared__ unsigned long long s_off, s_total;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const long long base = ((long long)blockIdx.x * PF_THREADS + (long long)warp * 32) * rows;
  if (src.skip()) return;  // (episodes: no update pending for this belief) - the same for every CTA

  // ---- phase A: propagate + weight
  double w[PF_ITEMS];
  unsigned alive_mask = 0;
  unsigned long long mx = 0, first = ~0ull;
#pragma unroll
  for (int j = 0; j < PF_ITEMS; ++j) {
    long long i = base + j * 32 + lane;
    w[j] = 0.0;
    if (j < rows && i < n) {
      typename Src::State sp;
      bool alive = src.eval(i, sp, w[j]);
      if (alive) {
        alive_mask |= 1u << j;
        if ((unsigned long long)i < first) first = (unsigned long long)i;
        unsigned long long b = (w[j] > 0.0) ? d2u(w[j]) : 0ull;
        mx = b > mx ? b : mx;
      } else {
        w[j] = 0.0;
      }
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    unsigned long long om = __shfl_xor_sync(0xffffffffu, mx, off), of = __shfl_xor_sync(0xffffffffu, first, off);
    mx = om > mx ? om : mx;
    first = of < first ? of : first;
  }
  if (lane == 0) { s_a[warp] = mx; s_b[warp] = first; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < PF_THREADS / 32; ++k) {
      mx = s_a[k] > mx ? s_a[k] : mx;
      first = s_b[k] < first ? s_b[k] : first;
    }
    if (first != ~0ull) {
      atomicMax(&ctl->max_bits, mx);
      atomicMax(&ctl->first_alive_inv, ~first);
      ctl->any_alive = 1;
    }
  }
  pf_grid_barrier(bar, gridDim.x);

  // ---- phase B: quantise and scan inside the CTA
  const unsigned long long max_bits = ld_volatile_u64(&ctl->max_bits);
  const int scale = pf_scale(max_bits, n);
  unsigned long long inc[PF_ITEMS];
  unsigned long long carry = 0;
#pragma unroll
  for (int j = 0; j < PF_ITEMS; ++j) {
    unsigned long long x = ((alive_mask >> j) & 1u) ? pf_quantise(w[j], scale) : 0ull;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += t;
    }
    inc[j] = x + carry;
    carry += __shfl_sync(0xffffffffu, x, 31);
  }
  if (lane == 0) s_a[warp] = carry;
  __syncthreads();
  if (warp == 0) {
    unsigned long long v = (lane < PF_THREADS / 32) ? s_a[lane] : 0ull, x = v;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += t;
    }
    if (lane < PF_THREADS / 32) s_a[lane] = x - v;  // exclusive over warps
    if (lane == 31) atomicExch(&cta_tot[blockIdx.x], x);
  }
  pf_grid_barrier(bar, 2u * gridDim.x);

  // ---- offsets: sum of the CTA totals before this CTA, and the grand total T
  pf_cta_offsets(cta_tot, s_b, s_c, s_off, s_total);
  const unsigned long long T = s_total;
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->total = T;
    int st = POMCP_OK;
    if (!ctl->any_alive) st = POMCP_ERR_ALL_SAMPLES_TERMINAL;
    else if (max_bits == 0 || T == 0) st = POMCP_ERR_PARTICLE_DEPLETION;
    ctl->status = st;
  }
  if (T == 0) return;

  // ---- phase C: particle i owns outputs [M(C_{i-1}), M(C_i))
  const unsigned long long first_alive = ~ld_volatile_u64(&ctl->first_alive_inv);
  const unsigned long long off0 = s_off + s_a[warp];
  const PfDiv dv(T, ru, n_out);
  long long prev_hi = dv.points_below(off0);  // M at the start of this warp's range
#pragma unroll
  for (int j = 0; j < PF_ITEMS; ++j) {
    if (j >= rows) break;  // uniform
    long long i = base + j * 32 + lane;
    long long hi = dv.points_below(off0 + inc[j]);
    long long lo = __shfl_up_sync(0xffffffffu, hi, 1);
    if (lane == 0) lo = prev_hi;
    prev_hi = __shfl_sync(0xffffffffu, hi, 31);
    if (j < rows && i < n && ((alive_mask >> j) & 1u)) {
      if ((unsigned long long)i == first_alive) lo = 0;
      if (lo < hi) {
        typename Src::State sp;
        src.propagate(i, sp);
        for (long long mth = lo; mth < hi; ++mth) pf_st_state(out + mth, (Out)sp);
      }
    }
  }
}

#ifdef POMCP_PROBE  // developer probe: %globaltimer of every CTA at start / barrier arrival / barrier exit / end
__device__ unsigned long long g_pf_probe[PF_MAX_GRID * 4];
__device__ __forceinline__ void pf_probe(int slot) {
  if (threadIdx.x == 0) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    g_pf_probe[blockIdx.x * 4 + slot] = t;
  }
}
#else
__device__ __forceinline__ void pf_probe(int) {}
#endif

// ---- K6+K7+K8 with ONE grid barrier: the a-priori scale (pf_scale_ceil) needs no max-weight pass, so phase A
// goes straight from the weights to the quantised CTA totals; the only grid-wide dependency left is the total T
// and the offsets of the CTAs, which every CTA derives from the <= 1024 CTA totals after the barrier.  When the
// a-priori scale does not fit these weights the kernel says so (PF_RETRY_MAX_SCALE) and writes nothing.
// BLOCKED: a thread owns `items` consecutive particles (states of <= 8 bytes: 16/32-byte vector-friendly, kept in
// registers for the emit); otherwise particles are striped over the warp (particle = base + row * 32 + lane), the
// only layout in which 16-byte states coalesce.
template <class Src, class Out, bool BLOCKED>
__global__ void __launch_bounds__(PF_THREADS, 2) k_pf_fused_ceil(const __grid_constant__ Src src, long long n, PfCtl* ctl,
                                                                 uint32_t ru, long long n_out, Out* out,
                                                                 unsigned long long* cta_tot, unsigned int* bar,
                                                                 int items, unsigned long long* next_ctl) {
  using State = typename Src::State;
  __shared__ unsigned long long s_a[PF_THREADS / 32];
  __shared__ unsigned long long s_b[PF_THREADS / 32];
  __shared__ unsigned long long s_c[PF_THREADS / 32];
  __shared__ unsigned long long s_off, s_total;
  // weight classes (discrete problems, blocked layout): quantised weights by table look-up, see ModelSrc::NCLASS
  constexpr bool CLASSES = Src::NCLASS > 0 && BLOCKED && sizeof(State) <= 8;
  constexpr bool CLASS_IN = CLASSES && Src::CLASS_IN;
  __shared__ unsigned long long qtab[CLASSES ? Src::NCLASS : 1];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  // the other control block, for the next launch (nobody uses it now)
  if (blockIdx.x == gridDim.x - 1)
    for (int k = threadIdx.x; k < PF_CTRL_WORDS; k += PF_THREADS) next_ctl[k] = 0ull;
  if (src.skip()) return;
  pf_probe(0);
  if (threadIdx.x == 0) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(ctl));
    asm volatile("prefetch.global.L2 [%0];" ::"l"(cta_tot + blockIdx.x));
  }
  const int scale = pf_scale_ceil(Src::CEIL_EXP, n);
  const long long base = BLOCKED ? ((long long)blockIdx.x * PF_THREADS + threadIdx.x) * items
                                 : ((long long)blockIdx.x * PF_THREADS + (long long)warp * 32) * items;
  auto index = [&](int k) { return BLOCKED ? base + k : base + (long long)k * 32 + lane; };

  // ---- phase A: propagate, weight, quantise.  States of <= 8 bytes stay in registers for the emit.
  constexpr bool KEEP = BLOCKED && sizeof(State) <= 8;
  State st[KEEP ? PF_ITEMS : 1];
  if (KEEP) {
#pragma unroll
    for (int k = 0; k < PF_ITEMS; ++k)
      if (k < items && index(k) < n) st[KEEP ? k : 0] = src.load(index(k));
  }
  unsigned long long q[PF_ITEMS];
  unsigned alive_mask = 0;
  unsigned long long first = ~0ull;
  bool over = false;
  // small sets go without the table: building it costs a round trip to the model's tables (2 us of a 10 us update)
  const bool cls = CLASSES && n >= 65536;  // uniform
  if constexpr (CLASSES) {
    if (cls) over = pf_fill_qtab(src, scale, qtab);  // (the state loads above are in flight meanwhile)
  }
  auto eval_direct = [&](int k, long long i) {
    State sp;
    double w;
    const bool alive = KEEP ? src.eval_from(i, st[KEEP ? k : 0], sp, w) : src.eval(i, sp, w);
    if (KEEP) st[KEEP ? k : 0] = sp;  // the propagated state is what the emit copies
    if (alive) {
      alive_mask |= 1u << k;
      if ((unsigned long long)i < first) first = (unsigned long long)i;
      over = over || pf_over_ceil(w, Src::CEIL_EXP);
      q[k] = pf_quantise(w, scale);
    }
  };
#pragma unroll
  for (int k = 0; k < PF_ITEMS; ++k) {
    q[k] = 0ull;
    const long long i = index(k);
    if (k < items && i < n) {
      if constexpr (CLASSES) {
        if (cls) {  // selects, not branches: the look-ups of a thread's particles overlap
          const State s0 = st[KEEP ? k : 0];
          const bool alive = !Src::M::terminal(src.m, s0);
          alive_mask |= alive ? 1u << k : 0u;
          first = alive && (unsigned long long)i < first ? (unsigned long long)i : first;
          if constexpr (CLASS_IN) {
            const unsigned long long qk = qtab[src.wclass(s0)];  // (in range for any bit pattern)
            q[k] = alive ? qk : 0ull;  // st keeps the INPUT state: the emit propagates the particles it copies
          } else {
            State sp;
            src.propagate_from(i, s0, sp);
            const unsigned long long qk = qtab[src.wclass(sp)];
            st[KEEP ? k : 0] = alive ? sp : s0;
            q[k] = alive ? qk : 0ull;
          }
        } else {
          eval_direct(k, i);
        }
      } else {
        eval_direct(k, i);
      }
    }
  }
  // inclusive cumulative sums in particle order inside the warp's range
  unsigned long long inc[PF_ITEMS];
  unsigned long long thread_excl = 0, warp_total;
  if (BLOCKED) {
    unsigned long long run = 0;
#pragma unroll
    for (int k = 0; k < PF_ITEMS; ++k) { run += q[k]; inc[k] = run; }
    unsigned long long x = run;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
      if (lane >= off) x += t;
    }
    thread_excl = x - run;
    warp_total = __shfl_sync(0xffffffffu, x, 31);
  } else {
    unsigned long long carry = 0;
#pragma unroll
    for (int k = 0; k < PF_ITEMS; ++k) {
      unsigned long long x = q[k];
#pragma unroll
      for (int off = 1; off < 32; off <<= 1) {
        unsigned long long t = __shfl_up_sync(0xffffffffu, x, off);
        if (lane >= off) x += t;
      }
      inc[k] = x + carry;
      carry += __shfl_sync(0xffffffffu, x, 31);
    }
    warp_total = carry;
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    unsigned long long of = __shfl_xor_sync(0xffffffffu, first, off);
    first = of < first ? of : first;
  }
  if (__any_sync(0xffffffffu, over) && lane == 0) atomicOr(&ctl->over_ceil, 1u);
  if (lane == 0) { s_a[warp] = warp_total; s_b[warp] = first; }
  __syncthreads();
  if (warp == 0) {
    unsigned long long v = (lane < PF_THREADS / 32) ? s_a[lane] : 0ull, y = v;
    unsigned long long f = (lane < PF_THREADS / 32) ? s_b[lane] : ~0ull;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      unsigned long long t = __shfl_up_sync(0xffffffffu, y, off);
      if (lane >= off) y += t;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      unsigned long long of = __shfl_xor_sync(0xffffffffu, f, off);
      f = of < f ? of : f;
    }
    if (lane < PF_THREADS / 32) s_a[lane] = y - v;  // exclusive over warps
    if (lane == 31) atomicExch(&cta_tot[blockIdx.x], y);
    if (lane == 0 && f != ~0ull) {
      atomicMax(&ctl->first_alive_inv, ~f);
      ctl->any_alive = 1;
    }
  }
  pf_probe(1);
  pf_grid_barrier(bar, gridDim.x);
  pf_probe(2);

  // ---- offsets: sum of the CTA totals before this CTA, and the grand total T
  pf_cta_offsets(cta_tot, s_b, s_c, s_off, s_total);
  const unsigned long long T = s_total;
  unsigned int over_any;
  asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(over_any) : "l"(&ctl->over_ceil));
  const bool retry = over_any != 0u || T < PF_MIN_TOTAL_CEIL;  // the same for every CTA
  if (blockIdx.x == 0 && threadIdx.x == 0) {
    ctl->total = T;
    ctl->status = retry ? PF_RETRY_MAX_SCALE : POMCP_OK;
  }
  if (retry) return;

  // ---- emit: particle i owns outputs [M(C_{i-1}), M(C_i))
  const unsigned long long first_alive = ~ld_volatile_u64(&ctl->first_alive_inv);
  const PfDiv dv(T, ru, n_out);
  if (BLOCKED) {
    const unsigned long long off0 = s_off + s_a[warp] + thread_excl;  // C just before this thread's particles
    const PfDiv::Fast fs = dv.fast(ru);
    const bool same = src.identity();  // uniform: the action leaves the states as they are
    long long lo = dv.points_below_fast2(off0, fs);
#pragma unroll
    for (int k = 0; k < PF_ITEMS; ++k) {
      const long long i = base + k;
      const long long hi = ((alive_mask >> k) & 1u) ? dv.points_below_fast2(off0 + inc[k], fs) : lo;
      if ((alive_mask >> k) & 1u) {
        if ((unsigned long long)i == first_alive) lo = 0;
        const uint32_t copies = (uint32_t)(hi - lo);  // n_out < 2^32
        if (copies) {
          State sp;
          if (CLASS_IN && cls && !same) src.propagate_from(i, st[KEEP ? k : 0], sp);
          else if (KEEP) sp = st[KEEP ? k : 0];
          else src.propagate(i, sp);
          Out* o = out + lo;
          const Out v = (Out)sp;
          pf_st_state(o, v);
          if (copies > 1) pf_st_state(o + 1, v);
          if (copies > 2) pf_st_state(o + 2, v);
          if (copies > 3) pf_st_state(o + 3, v);
          if (copies > 4)
            for (uint32_t j = 4; j < copies; ++j) pf_st_state(o + j, v);
        }
      }
      lo = hi;
    }
  } else {
    const unsigned long long off0 = s_off + s_a[warp];
    const double ru_frac = (double)ru * 2.3283064365386963e-10;  // ru / 2^32, exact
    long long prev_hi = dv.points_below_fast(off0, ru_frac);  // M at the start of this warp's range
#pragma unroll
    for (int k = 0; k < PF_ITEMS; ++k) {
      if (k >= items) break;  // uniform
      const long long i = base + (long long)k * 32 + lane;
      const long long hi = dv.points_below_fast(off0 + inc[k], ru_frac);
      long long lo = __shfl_up_sync(0xffffffffu, hi, 1);
      if (lane == 0) lo = prev_hi;
      prev_hi = __shfl_sync(0xffffffffu, hi, 31);
      if (i < n && ((alive_mask >> k) & 1u)) {
        if ((unsigned long long)i == first_alive) lo = 0;
        if (lo < hi) {
          State sp;
          src.propagate(i, sp);
          for (long long mth = lo; mth < hi; ++mth) pf_st_state(out + mth, (Out)sp);
        }
      }
    }
  }
  __syncthreads();
  pf_probe(3);
}

// n draws from the problem's initial state distribution (the reinvigorator's stand-in for an old
// belief that is still the distribution itself)
template <int PID>
__global__ void k_init_particles(const __grid_constant__ DevModel m, typename ModelT<PID>::State* out, long long n,
                                 uint32_t k0, uint32_t k1) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  Philox4 p = philox4x32_10(0u, TAG_INIT, (uint32_t)i, (uint32_t)((unsigned long long)i >> 32), k0, k1);
  out[i] = ModelT<PID>::initial(m, u01(p.w[0]), u01(p.w[1]));
}

// generate_sor for the outer simulation loop (POMDPs.generate_sor): one environment
// step of the problem model, drawn from the Philox stream (seed; counter (0, TAG_ENV, step)).
template <int PID>
struct EnvStep {
  typename ModelT<PID>::State sp;
  unsigned long long obs;
  double r;
  int terminal;
};
template <int PID>
__global__ void k_generate_sor(const __grid_constant__ DevModel m, const typename ModelT<PID>::State* s_in, int a,
                               uint32_t k0, uint32_t k1, uint32_t step, EnvStep<PID>* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  using M = ModelT<PID>;
  Philox4 p = philox4x32_10(0u, TAG_ENV, step, 0u, k0, k1);
  uint64_t o;
  M::template step<true, double>(m, s_in[0], a, u01(p.w[0]), u01(p.w[2]), u01(p.w[3]), out->sp, o, out->r);
  out->obs = o;
  out->terminal = M::terminal(m, out->sp) ? 1 : 0;
}
template <int PID>
__global__ void k_initial_state(const __grid_constant__ DevModel m, uint32_t k0, uint32_t k1,
                                typename ModelT<PID>::State* out) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  Philox4 p = philox4x32_10(1u, TAG_ENV, 0u, 0u, k0, k1);
  out[0] = ModelT<PID>::initial(m, u01(p.w[0]), u01(p.w[1]));
}

// pomcp_pf_weights parity hook
template <int PID>
__global__ void k_pf_weights(const __grid_constant__ ModelSrc<PID> src, long long n, double* w_out,
                             typename ModelT<PID>::State* sp_out) {
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  typename ModelT<PID>::State sp;
  double w;
  src.eval(i, sp, w);
  w_out[i] = w;
  sp_out[i] = sp;
}

// ---- K9: gather the particles pushed at one belief node (stable stream compaction of the
// append-only log, fused into the scan's sink).
struct LogMatchSrc {
  const uint32_t* log_node;
  uint32_t target;
  __device__ __forceinline__ unsigned long long operator()(long long i) const {
    return log_node[i] == target ? 1ull : 0ull;
  }
};
template <class State>
struct LogGatherSink {
  const State* log_state;
  State* out;
  long long cap;
  __device__ __forceinline__ void operator()(long long i, unsigned long long inc, unsigned long long v) const {
    if (v && (long long)inc <= cap) out[inc - 1] = log_state[i];
  }
};
template <class State>
__global__ void __launch_bounds__(SCAN_THREADS) k_log_gather(const uint32_t* log_node, const State* log_state,
                                                             uint32_t target, long long n, State* out,
                                                             long long cap, unsigned long long* status,
                                                             unsigned long long* ticket) {
  LogMatchSrc s{log_node, target};
  LogGatherSink<State> k{log_state, out, cap};
  chained_scan(s, k, n, status, ticket);
}
__global__ void k_scan_total(const unsigned long long* status, int n_tiles, unsigned long long* total_out) {
  if (threadIdx.x == 0 && blockIdx.x == 0) *total_out = status[n_tiles - 1] & SCAN_VAL_MASK;
}

// pomcp_rollout_batch parity hook (K3 alone)
template <int PID>
__global__ void __launch_bounds__(256) k_rollout_batch(const __grid_constant__ DevModel m,
                                                       const typename ModelT<PID>::State* states, long long n,
                                                       int steps, uint32_t k0, uint32_t k1, double* ret, int* ns) {
  __shared__ RsTab rs_tab;
  if (PID == POMCP_ROCKSAMPLE) rs_tab_init(&rs_tab, m);
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (i >= n) return;
  int nsteps = 0;
  ret[i] = rollout_random<PID>(m, &rs_tab, states[i], steps, k0, k1, TAG_ROLLOUT, (uint32_t)i, 0u, nsteps);
  ns[i] = nsteps;
}

__global__ void k_flush_l2(float4* buf, size_t n) {
  size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  size_t stride = (size_t)gridDim.x * blockDim.x;
  for (; i < n; i += stride) buf[i] = make_float4(1.f, 2.f, 3.f, 4.f);
}

}  // namespace pomcp

// K0 — counter-based Philox4x32-10 (Salmon et al., SC'11) for device and host.
// Replaces the shared mutable `sol.rng::AbstractRNG` threaded through the reference
// (src/solver.jl:48,126; src/rollout.jl:79): every draw is addressed by
// counter = (index, tag | rank<<8, query, epoch) under key = 64-bit seed, so thousands of
// simulations can draw independently and the CPU oracle can reproduce every word.
#pragma once
#include <cstdint>

#ifdef __CUDACC__
#define PX_HD __host__ __device__ __forceinline__
#else
#define PX_HD inline
#endif

namespace pomcp {

enum : uint32_t { TAG_ROOT = 0, TAG_TREE = 1, TAG_ROLLOUT = 2, TAG_PF = 3, TAG_RESAMPLE = 4, TAG_INIT = 5, TAG_ENV = 6 };

struct Philox4 {
  uint32_t w[4];
};

PX_HD void mulhilo32(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  lo = a * b;
  hi = __umulhi(a, b);
#else
  uint64_t p = (uint64_t)a * b;
  lo = (uint32_t)p;
  hi = (uint32_t)(p >> 32);
#endif
}

PX_HD Philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    mulhilo32(0xD2511F53u, c0, hi0, lo0);
    mulhilo32(0xCD9E8D57u, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
  Philox4 out;
  out.w[0] = c0; out.w[1] = c1; out.w[2] = c2; out.w[3] = c3;
  return out;
}

// uniform in [0,1) with 32-bit resolution: word * 2^-32 (exact in fp64)
PX_HD double u01(uint32_t w) { return (double)w * 2.3283064365386963e-10; }

}  // namespace pomcp

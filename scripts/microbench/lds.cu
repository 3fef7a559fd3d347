// Microbenchmark: latency of the dependent chain of a rollout step (LOP3 -> shared-memory load -> LOP3 ...), with the
// bank conflicts of the transition table's layout (mode 1) and without (mode 0 / 2).  DESIGN.md 4.1: 37 cycles per step.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o lds lds.cu && ./lds
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
// pointer chase through shared memory: off = (e & 0x3fc3) | (a4 & 0x3c); e = tab[off]
__global__ void k(int mode, int iters, uint32_t* out, long long* cyc) {
  __shared__ uint32_t tab[4096];
  for (int i = threadIdx.x; i < 4096; i += blockDim.x) {
    int row = i >> 4, a = i & 15;
    int x = row & 15, y = row >> 4;
    uint32_t next;
    if (mode == 0) next = row;                  // stay: every lane keeps its row
    else { // random-walk-ish: deterministic pseudo-random next among the 7x7 cells
      uint32_t h = (uint32_t)i * 2654435761u; h ^= h >> 13;
      next = ((h >> 3) % 7) * 16 + (h % 7);
    }
    tab[i] = next << 6;
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  uint32_t e = (mode == 2 ? 0u : (uint32_t)((lane % 7) * 16 + (lane / 7) % 7)) << 6;
  uint32_t rng = lane * 747796405u + 12345u;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
    rng = rng * 1664525u + 1013904223u;
    uint32_t a4 = mode == 2 ? 0u : __umulhi(rng, 52u);
    uint32_t off;
    asm volatile("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(off) : "r"(e), "r"(a4), "n"(0x3fc3));
    e = *(volatile uint32_t*)((char*)tab + off);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = e;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  uint32_t* out; long long* cyc;
  cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 1024 * 8);
  for (int mode = 0; mode < 3; ++mode)
    for (int warps : {1, 4, 16}) {
      for (int rep = 0; rep < 2; ++rep) k<<<1, 32 * warps>>>(mode, 10000, out, cyc);
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      printf("mode %d warps/CTA %d: %.1f cycles per dependent LDS step\n", mode, warps, c / 10000.0);
    }
  return 0;
}

// Microbenchmark of the RockSample rollout loop as the search kernel compiles it: cycles per step
// of one warp alone on an SM, and with 4 / 8 / 16 warps per SM, one (NT 1) or two (NT 2) trajectories per lane.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o roll roll.cu && ./roll
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cuda_runtime.h>
#include "../../pomcp.jl_b200/csrc/tree.cuh"
using namespace pomcp;
__global__ void k(const __grid_constant__ DevModel m, int steps, int reps, int nt, double* out, long long* cyc, int* ns) {
  __shared__ RsTab tb;

  rs_tab_init(&tb, m);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  double acc = 0; int nst = 0;
  long long t0 = clock64();
  for (int r = 0; r < reps; ++r) {
    int n = 0;
    uint32_t s = m.rs_start | (0xA5u << 8);
    if (nt == 1) {
      acc += rollout_random<POMCP_ROCKSAMPLE, false>(m, &tb, s, steps, 1u, 2u, TAG_ROLLOUT | ((uint32_t)(lane + 32 * warp) << 16), (uint32_t)(r + 1000 * blockIdx.x), 0u, n);
    } else {
      const uint32_t tw[2] = {TAG_ROLLOUT | ((uint32_t)(lane + 64 * warp) << 16), TAG_ROLLOUT | ((uint32_t)(lane + 32 + 64 * warp) << 16)};
      const bool en[2] = {true, true};
      double tot[2];
      rollout_rocksample<false, 2>(m, &tb, s, steps, 1u, 2u, tw, en, (uint32_t)(r + 1000 * blockIdx.x), 0u, n, tot);
      acc += tot[0] + tot[1];
    }
    nst += n;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  ns[blockIdx.x * blockDim.x + threadIdx.x] = nst;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
  DevModel m; memset(&m, 0, sizeof m);
  m.pid = POMCP_ROCKSAMPLE; m.nA = 13; m.nO = 3; m.state_bytes = 4; m.gamma = 0.95;
  m.rs_n = 7; m.rs_k = 8; m.rs_start = 0 | (3 << 4);
  m.rs_rtab[0] = 0; m.rs_rtab[1] = 10; m.rs_rtab[2] = -10; m.rs_rtab[3] = 10;
  memset(m.rock_at, -1, 256);
  int rx[8] = {2, 0, 3, 6, 2, 3, 5, 1}, ry[8] = {0, 1, 1, 3, 4, 4, 5, 6};
  for (int i = 0; i < 8; ++i) m.rock_at[ry[i] * 16 + rx[i]] = (signed char)i;
  double* out; long long* cyc; int* ns;
  cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 8192); cudaMalloc(&ns, 1 << 22);
  for (int nt : {1, 2}) for (int warps : {1, 4, 8, 16}) {
    const int reps = 200, steps = 85;
    for (int it = 0; it < 2; ++it) k<<<1, 32 * warps>>>(m, steps, reps, nt, out, cyc, ns);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    static int h[512]; cudaMemcpy(h, ns, 4 * 32 * warps, cudaMemcpyDeviceToHost);
    long long live = 0; for (int i = 0; i < 32 * warps; ++i) live += h[i];
    printf("NT %d warps %2d on one SM: %.0f cycles per 85-step rollout call (%.1f per executed step), live steps per lane-call %.1f  err=%d\n", nt, warps,
           c / (double)reps, c / (double)reps / steps, live / (double)(32 * warps * reps), (int)cudaGetLastError());
  }
  return 0;
}

// Exhaustive check of the table-driven RockSample tree step (rs_tree_step) against ModelT<ROCKSAMPLE>::step<true>: every
// cell, rock set and action of RockSample(7,8), eight sensor draws each (1.3e6 cases).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -fmad=false -o treestep treestep.cu && ./treestep
#include <cstdio>
#include <cstdint>
#include <cstring>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>
#include "../../pomcp.jl_b200/csrc/tree.cuh"
using namespace pomcp;
__global__ void k(const __grid_constant__ DevModel m, const uint32_t* eta_thr, int* bad, uint32_t* ex) {
  __shared__ RsTab tb;
  rs_tab_init(&tb, m);
  const int nA = m.nA;
  for (uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x; idx < 49u * 256u * (uint32_t)nA * 8u; idx += gridDim.x * blockDim.x) {
    uint32_t t = idx;
    const int wi = t % 8; t /= 8;
    const int a = t % nA; t /= nA;
    const uint32_t rocks = t % 256; t /= 256;
    const int x = t % 7, y = t / 7;
    const uint32_t st = (uint32_t)x | ((uint32_t)y << 4) | (rocks << 8);
    const uint32_t w = (uint32_t)wi * 0x24924925u + 12345u * (idx % 977);
    uint32_t sp1, sp2; uint64_t o1, o2; double r1, r2;
    rs_tree_step(m, &tb, eta_thr, st, a, w, sp1, o1, r1);
    ModelT<POMCP_ROCKSAMPLE>::step<true, double>(m, st, a, 0.0, u01(w), 0.0, sp2, o2, r2);
    if (sp1 != sp2 || o1 != o2 || r1 != r2) { if (atomicAdd(bad, 1) == 0) { ex[0] = st; ex[1] = a; ex[2] = sp1; ex[3] = sp2; ex[4] = (uint32_t)o1; ex[5] = (uint32_t)o2; } }
  }
}
int main() {
  DevModel m; memset(&m, 0, sizeof m);
  m.pid = POMCP_ROCKSAMPLE; m.nA = 13; m.nO = 3; m.state_bytes = 4; m.gamma = 0.95;
  m.rs_n = 7; m.rs_k = 8; m.rs_start = 0 | (3 << 4);
  m.rs_r_good = 10; m.rs_r_bad = -10; m.rs_r_exit = 10;
  m.rs_rtab[0] = 0; m.rs_rtab[1] = 10; m.rs_rtab[2] = -10; m.rs_rtab[3] = 10;
  memset(m.rock_at, -1, 256);
  int rx[8] = {2, 0, 3, 6, 2, 3, 5, 1}, ry[8] = {0, 1, 1, 3, 4, 4, 5, 6};
  for (int i = 0; i < 8; ++i) m.rock_at[ry[i] * 16 + rx[i]] = (signed char)i;
  std::vector<double> eta(256 * POMCP_MAX_ROCKS, 0.0);
  std::vector<uint32_t> thr(7 * 7 * 8);
  for (int y = 0; y < 7; ++y) for (int x = 0; x < 7; ++x) for (int i = 0; i < 8; ++i) {
    double d = std::sqrt((double)((x - rx[i]) * (x - rx[i]) + (y - ry[i]) * (y - ry[i])));
    double e = 0.5 * (1.0 + std::exp2(-d / 20.0));
    eta[(size_t)(y * 16 + x) * POMCP_MAX_ROCKS + i] = e;
    thr[(size_t)(y * 7 + x) * 8 + i] = (uint32_t)(std::ceil(e * 4294967296.0) - 1.0);
  }
  double* d_eta; uint32_t* d_thr; int* bad; uint32_t* ex;
  cudaMalloc(&d_eta, eta.size() * 8); cudaMemcpy(d_eta, eta.data(), eta.size() * 8, cudaMemcpyHostToDevice);
  cudaMalloc(&d_thr, thr.size() * 4); cudaMemcpy(d_thr, thr.data(), thr.size() * 4, cudaMemcpyHostToDevice);
  cudaMalloc(&bad, 4); cudaMemset(bad, 0, 4); cudaMalloc(&ex, 32);
  m.rs_eta = d_eta;
  k<<<148, 256>>>(m, d_thr, bad, ex);
  int hb; uint32_t hex[6]; cudaMemcpy(&hb, bad, 4, cudaMemcpyDeviceToHost); cudaMemcpy(hex, ex, 24, cudaMemcpyDeviceToHost);
  printf("mismatches %d of %d  err=%d  first: st=%x a=%u sp=%x/%x o=%u/%u\n", hb, 49 * 256 * 13 * 8, (int)cudaGetLastError(), hex[0], hex[1], hex[2], hex[3], hex[4], hex[5]);
  return 0;
}

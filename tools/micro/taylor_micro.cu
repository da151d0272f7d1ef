// Standalone timing harness for taylor2_kernel variants (compile-time experiment macros).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 [-DT2_EXP_...] -o taylor_micro taylor_micro.cu
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <algorithm>
#include "../../pauxy_b200/csrc/pxb_taylor2.cuh"
using namespace pxb;
#ifndef EXP_NG
#define EXP_NG 2
#endif
#ifndef EXP_WMX
#define EXP_WMX 4
#endif
#ifndef EXP_WNX
#define EXP_WNX 6
#endif
int main(int argc, char** argv) {
  int W = argc > 1 ? atoi(argv[1]) : 2368;
  int dbg = argc > 2 ? atoi(argv[2]) : 1;
  int nbuf = argc > 3 ? atoi(argv[3]) : 2;
  Dims d{};
  d.M = 108; d.na = d.nb = 21; d.ne = 42; d.N = 500; d.W = W; d.Wtot = W;
  d.Mp = 108; d.KC = 27; d.M8 = 112; d.MT = 14; d.Wp = W; d.WG = W / 4; d.exp_order = 6;
  size_t vf = (size_t)W * vf_walker(d), of = of_size(d);
  double *VF, *phi;
  cudaMalloc(&VF, vf * 8); cudaMalloc(&phi, of * 8);
  std::vector<double> h(1 << 20);
  for (auto& x : h) x = (rand() / (double)RAND_MAX - 0.5) * 0.02;
  for (size_t o = 0; o < vf; o += h.size()) cudaMemcpy(VF + o, h.data(), std::min(h.size(), vf - o) * 8, cudaMemcpyHostToDevice);
  for (size_t o = 0; o < of; o += h.size()) cudaMemcpy(phi + o, h.data(), std::min(h.size(), of - o) * 8, cudaMemcpyHostToDevice);
  Taylor2Args a{};
  a.VF = VF; a.phi = phi; a.active = nullptr; a.d = d; a.ochunk = 44; a.nchunks = 1; a.NT = 11; a.dbg = dbg;
  const int NG = EXP_NG;
  int msize[4], nsize[4] = {0, 0, 0, 0};
  a.m_off[0] = 0;
  for (int g = 0; g < 4; ++g) { msize[g] = d.MT / 4 + (g < d.MT % 4); a.m_off[g + 1] = a.m_off[g] + msize[g]; }
  a.n_off[0] = 0;
  for (int g = 0; g < NG; ++g) { nsize[g] = a.NT / NG + (g < a.NT % NG); a.n_off[g + 1] = a.n_off[g] + nsize[g]; }
  for (int g = NG; g < 4; ++g) a.n_off[g + 1] = a.n_off[NG];
  int load[4] = {0, 0, 0, 0};
  for (int g = 0; g < 4; ++g) {
    int order[4] = {0, 1, 2, 3};
    std::sort(order, order + 4, [&](int x, int y) { return load[x] != load[y] ? load[x] < load[y] : x < y; });
    for (int k = 0; k < 4; ++k) { a.mperm[g][order[k]] = k; if (g < NG) load[order[k]] += msize[k] * nsize[g]; }
  }
  a.nbuf = nbuf;
  a.nstage = 0;
  for (int ns = 12; ns >= 2; --ns) if (taylor2_smem_bytes(d, a.NT, nbuf, ns) <= 232448) { a.nstage = ns; break; }
  size_t smem = taylor2_smem_bytes(d, a.NT, a.nbuf, a.nstage);
  auto kern = taylor2_kernel<EXP_WMX, EXP_WNX, EXP_NG>;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  float best = 1e9;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    kern<<<148, T2Cfg<EXP_NG>::threads, smem>>>(a);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); best = std::min(best, ms);
  }
  cudaError_t err = cudaGetLastError();
  double dmma = (double)W * 14 * 11 * 27 * 2 * 6;
  printf("W=%d dbg=%d nbuf=%d nstage=%d KS=%d NG=%d: %.3f ms, %.2f TFLOP/s executed (%.1f%% of 37.1), err=%s\n", W, dbg, nbuf, a.nstage,
         T2_KS, NG, best, dmma * 512 / best * 1e-9, dmma * 512 / best * 1e-9 / 37.1 * 100, cudaGetErrorString(err));
  return 0;
}

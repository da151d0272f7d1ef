// Micro-benchmark of the consumer main loop of the Taylor kernels in isolation: N warps per SM, each
// owning a WM x WN block of tile pairs, fragments read from (static) shared memory, no ring, no
// barriers, no epilogue.  Answers: which (product form, warp count, tile shape, register budget,
// loop style) keeps the FP64 DMMA pipe busy?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -o tm_dmma_loop dmma_loop_micro.cu
#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}
__device__ __forceinline__ double flip(double v, unsigned mask) {
  return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
}

constexpr int KC = 27, MT = 14, ORDERS = 6, NSTAGE = 4, KS = 2;

// MODE 4: interleaved complex (2 DMMAs per tile pair: ar*b, ai*(iB)); MODE 3: planar 3-product
// PIPE 0: load fragments, then DMMAs (one fragment set); PIPE 1: two fragment sets, k-steps unrolled x2,
// next set requested before the current DMMAs
template <int MODE, int WM, int WN, int PIPE>
__device__ __forceinline__ void body(const double* smA, const double* smB, int m0, int n0, int lane, int nslots,
                                     double* out, int walkers) {
  constexpr int NP = MODE == 3 ? 3 : 1;
  double acc[NP][WM][WN][2];
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) acc[p][i][j][0] = acc[p][i][j][1] = 0.0;
  const unsigned smask = (lane & 4) ? 0u : 0x80000000u;
  struct Frag { double ar[WM], ai[WM], b0[WN], b1[WN]; };
  auto load = [&](Frag& f, int kc) {
    const double* as = smA + ((kc / KS) % NSTAGE) * (MT * KS * 64) + (m0 * KS + (kc % KS)) * 64 + lane;
    const double* bs = smB + (size_t)kc * nslots * 32 + n0 * (MODE == 3 ? 64 : 32) + lane;
#pragma unroll
    for (int i = 0; i < WM; ++i) {
      f.ar[i] = as[i * KS * 64];
      f.ai[i] = as[i * KS * 64 + 32];
    }
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      if (MODE == 3) {
        f.b0[j] = bs[j * 64];
        f.b1[j] = bs[j * 64 + 32];
      } else {
        f.b0[j] = bs[j * 32];
        f.b1[j] = flip(bs[j * 32 + ((lane & 4) ? -4 : 4)], smask);
      }
    }
  };
  auto mma = [&](const Frag& f) {
    if (MODE == 3) {
      double as[WM], bs[WN];
#pragma unroll
      for (int i = 0; i < WM; ++i) as[i] = f.ar[i] + f.ai[i];
#pragma unroll
      for (int j = 0; j < WN; ++j) bs[j] = f.b0[j] + f.b1[j];
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma(acc[0][i][j][0], acc[0][i][j][1], f.ar[i], f.b0[j]);
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma(acc[1 % NP][i][j][0], acc[1 % NP][i][j][1], f.ai[i], f.b1[j]);
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma(acc[2 % NP][i][j][0], acc[2 % NP][i][j][1], as[i], bs[j]);
    } else {
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma(acc[0][i][j][0], acc[0][i][j][1], f.ar[i], f.b0[j]);
#pragma unroll
      for (int i = 0; i < WM; ++i)
#pragma unroll
        for (int j = 0; j < WN; ++j) dmma(acc[0][i][j][0], acc[0][i][j][1], f.ai[i], f.b1[j]);
    }
  };
  for (int w = 0; w < walkers; ++w) {
    for (int n = 0; n < ORDERS; ++n) {
      if (PIPE == 0) {
#pragma unroll 1
        for (int kc = 0; kc < KC; ++kc) {
          Frag f;
          load(f, kc);
          mma(f);
        }
      } else {
        Frag f0, f1;
        load(f0, 0);
#pragma unroll 1
        for (int kc = 0; kc + 1 < KC; kc += 2) {
          load(f1, kc + 1);
          mma(f0);
          load(f0, kc + 2 < KC ? kc + 2 : 0);
          mma(f1);
        }
        mma(f0);
      }
    }
  }
  double s = 0.0;
#pragma unroll
  for (int p = 0; p < NP; ++p)
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) s += acc[p][i][j][0] + acc[p][i][j][1];
  if (s == 1.2345) out[threadIdx.x] = s;
}

template <int MODE, int WM, int WN, int PIPE, int NWARPS>
__global__ void __launch_bounds__(NWARPS * 32, 1) loop_kernel(const double* init, double* out, int walkers, int nslots) {
  extern __shared__ __align__(128) double sm[];
  double* smA = sm;
  double* smB = sm + NSTAGE * MT * KS * 64;
  const int total = NSTAGE * MT * KS * 64 + KC * nslots * 32;
  for (int i = threadIdx.x; i < total; i += blockDim.x) sm[i] = init[i % 4096];
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m0 = ((warp % 4) * 3) % (MT - WM + 1);
  const int nmax = (MODE == 3 ? nslots / 2 : nslots) - WN;
  const int n0 = nmax > 0 ? (warp / 4) % (nmax + 1) : 0;
  body<MODE, WM, WN, PIPE>(smA, smB, m0, n0, lane, nslots, out, walkers);
}

template <int MODE, int WM, int WN, int PIPE, int NWARPS>
void run(const char* name, const double* init, double* out) {
  const int walkers = 16;
  const int nslots = 12;  // 32-double fragment slots per kc row of the iterate (12 x 256 B)
  auto kern = loop_kernel<MODE, WM, WN, PIPE, NWARPS>;
  const size_t smem = (size_t)(NSTAGE * MT * KS * 64 + KC * nslots * 32) * 8;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  cudaFuncAttributes fa;
  cudaFuncGetAttributes(&fa, kern);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  float best = 1e9f;
  for (int r = 0; r < 4; ++r) {
    cudaEventRecord(e0);
    kern<<<148, NWARPS * 32, smem>>>(init, out, walkers, nslots);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    best = std::min(best, ms);
  }
  const double per_k = (MODE == 3 ? 3.0 : 2.0) * WM * WN;
  const double dm = 148.0 * NWARPS * walkers * ORDERS * KC * per_k;
  const double rate = dm * 512 / best * 1e-9;
  printf("%-34s warps %2d tile %dx%d regs %3d spill %4zu: %7.3f ms  %6.2f TFLOP/s = %5.1f%% of 37.1   err=%s\n", name,
         NWARPS, WM, WN, fa.numRegs, (size_t)fa.localSizeBytes, best, rate, rate / 37.1 * 100,
         cudaGetErrorString(cudaGetLastError()));
}

int main() {
  double *init, *out;
  cudaMalloc(&init, 4096 * 8);
  cudaMalloc(&out, 4096 * 8);
  double h[4096];
  for (int i = 0; i < 4096; ++i) h[i] = (rand() / (double)RAND_MAX - 0.5) * 1e-3;
  cudaMemcpy(init, h, sizeof h, cudaMemcpyHostToDevice);
  run<4, 4, 3, 0, 16>("4M plain   (taylor2 NG=4 shape)", init, out);
  run<4, 4, 3, 1, 16>("4M pipe    (taylor2 NG=4 shape)", init, out);
  run<4, 4, 6, 0, 8>("4M plain   (8 warps 4x6)", init, out);
  run<4, 4, 6, 1, 8>("4M pipe    (8 warps 4x6)", init, out);
  run<4, 4, 4, 0, 12>("4M plain   (12 warps 4x4)", init, out);
  run<4, 4, 4, 1, 12>("4M pipe    (12 warps 4x4)", init, out);
  run<3, 4, 3, 0, 8>("3M plain   (8 warps 4x3)", init, out);
  run<3, 4, 3, 1, 8>("3M pipe    (8 warps 4x3)", init, out);
  run<3, 4, 2, 0, 12>("3M plain   (12 warps 4x2)", init, out);
  run<3, 4, 2, 1, 12>("3M pipe    (12 warps 4x2)", init, out);
  run<3, 2, 3, 0, 16>("3M plain   (16 warps 2x3)", init, out);
  run<3, 2, 3, 1, 16>("3M pipe    (16 warps 2x3)", init, out);
  run<3, 3, 2, 0, 16>("3M plain   (16 warps 3x2)", init, out);
  run<3, 4, 1, 0, 16>("3M plain   (16 warps 4x1)", init, out);
  run<3, 4, 1, 1, 16>("3M pipe    (16 warps 4x1)", init, out);
  return 0;
}

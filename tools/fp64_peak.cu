// FP64 pipe microbenchmark for sm_100a: register-resident DMMA.8x8x4 and DFMA
// issue rates, to fix the FP64 "tensor" roofline denominator beside cuBLAS DGEMM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_peak fp64_peak.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int NACC>
__global__ void k_dmma(double* out, int iters, double a0, double b0) {
  double c[NACC][2];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i][0] = c[i][1] = 0.0;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) dmma(c[i][0], c[i][1], a, b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int NACC>
__global__ void k_dfma(double* out, int iters, double a0, double b0) {
  double c[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) c[i] = i;
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) c[i] = fma(a, c[i], b);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

// DMMA and DFMA interleaved (are they separate pipes?)
template <int NACC>
__global__ void k_mixed(double* out, int iters, double a0, double b0) {
  double c[NACC][2], f[NACC];
#pragma unroll
  for (int i = 0; i < NACC; ++i) { c[i][0] = c[i][1] = 0.0; f[i] = i; }
  double a = a0 + threadIdx.x * 1e-9, b = b0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < NACC; ++i) {
      dmma(c[i][0], c[i][1], a, b);
      f[i] = fma(a, f[i], b);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < NACC; ++i) s += c[i][0] + c[i][1] + f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <class F>
float time_it(F f) {
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0);
  for (int r = 0; r < 5; ++r) f();
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  return ms / 5;
}

int main() {
  int dev = 0; cudaDeviceProp p; cudaGetDeviceProperties(&p, dev);
  int sms = p.multiProcessorCount;
  printf("device %s, %d SMs, clock %.0f MHz\n", p.name, sms, p.clockRate / 1e3);
  double* out; cudaMalloc(&out, sizeof(double) * sms * 32 * 1024);
  const int iters = 20000;
  for (int warps : {4, 8, 16, 32}) {
    int threads = warps * 32 > 1024 ? 1024 : warps * 32;
    int ctas = sms * (warps * 32 / threads);
    {
      float ms = time_it([&] { k_dmma<16><<<ctas, threads>>>(out, iters, 1.0, 1.0); });
      double flop = 2.0 * 256 * 16 * (double)iters * ctas * (threads / 32);
      printf("DMMA.884  nacc=16 warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s  (%.2f clk/DMMA/SM at 1.9GHz)\n", warps, ms,
             flop / ms * 1e-9, ms * 1e-3 * 1.9e9 / (16.0 * iters * (threads / 32) * (ctas / sms)));
    }
    {
      float ms = time_it([&] { k_dmma<4><<<ctas, threads>>>(out, iters, 1.0, 1.0); });
      double flop = 2.0 * 256 * 4 * (double)iters * ctas * (threads / 32);
      printf("DMMA.884  nacc= 4 warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
    }
    {
      float ms = time_it([&] { k_dfma<16><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
      double flop = 2.0 * 32 * 16 * (double)iters * ctas * (threads / 32);
      printf("DFMA      nacc=16 warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
    }
    {
      float ms = time_it([&] { k_mixed<8><<<ctas, threads>>>(out, iters, 1.0000001, 1e-9); });
      double flop = 2.0 * (256 + 32) * 8 * (double)iters * ctas * (threads / 32);
      printf("DMMA+DFMA nacc= 8 warps/SM=%2d : %8.3f ms  %7.2f TFLOP/s\n", warps, ms, flop / ms * 1e-9);
    }
  }
  cudaFree(out);
  return 0;
}

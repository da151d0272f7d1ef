// Fragment-major FP64 tensor-core GEMM:  C[mt][nt] = sum_ks A[mt][ks] * B[nt][ks]
// with A tiles 8x4 and B tiles 4x8 stored as 32 contiguous doubles per
// (tile, k-step).  One kernel, three epilogues:
//   EpiX    X_s[w][n]   = sum_k R_s[k,n] Theta_s[w,k]      (force bias / Coulomb; generic.py:130-152,
//                                                            estimators/generic.py:185-186)
//   EpiVHS  VHS_w[p][q] = i sqrt(dt) sum_n L[pq,n] x[w,n]   (generic.py:164-179) in Taylor A-operand order
//   EpiOF   phi'_s      = BH1_s phi_s                       (operations.py:29-52) back into OF layout
//
// Operands stream from L2/L1 with warp-contiguous 256 B loads, register
// double-buffered one k-step ahead; each warp owns a WM x WN block of 8x8
// tiles (WM*WN DMMA.8x8x4 per k-step for WM+WN loads).
#pragma once
#include "pxb_common.cuh"

namespace pxb {

struct GemmArgs {
  const double* A;     // [z][mt][KS][32]
  const double* B;     // tile nt at (nt / ntInner)*strideBO + (nt % ntInner)*strideBI (+ z*strideBz)
  size_t strideAz, strideBz, strideBO, strideBI;
  int ntInner;
  int MTiles, NTiles, KS;
};

struct EpiX {  // X[z][w][n] complex128, ld = Np
  double* X;
  int Wp, Np;
  __device__ __forceinline__ void operator()(int mt, int nt, int z, int g, int t, double c0,
                                             double c1) const {
    int w = nt * 4 + t, n = mt * 8 + g;
    double2* dst = reinterpret_cast<double2*>(X) + ((size_t)z * Wp + w) * Np + n;
    *dst = make_double2(c0, c1);
  }
};

struct EpiVHS {  // VF[w][MT][KC][c][g][t]; row tile mt = (mtv*4+s)*KC + kc
  double* VF;
  int KC, MT;
  double sqrt_dt;
  size_t walker_stride;
  __device__ __forceinline__ void operator()(int rt, int nt, int z, int g, int t, double c0,
                                             double c1) const {
    int kc = rt % KC, ms = rt / KC, s = ms & 3, mtv = ms >> 2;
    int w = nt * 4 + t;
    double* base = VF + (size_t)w * walker_stride + ((size_t)mtv * KC + kc) * 64 + 8 * s + g;
    // VHS = i sqrt(dt) (S_re + i S_im)
    base[0] = -sqrt_dt * c1;
    base[32] = sqrt_dt * c0;
  }
};

struct EpiOF {  // out OF buffer; n-tile nt = wg*nInner + il, orbital i = ioff + il
  double* out;
  const int* active;  // optional per-walker mask (skip store if 0)
  int ne, KC, ioff, nInner;
  __device__ __forceinline__ void operator()(int mt, int nt, int z, int g, int t, double c0,
                                             double c1) const {
    int wg = nt / nInner, il = nt % nInner;
    int pc = 2 * mt + (g >> 2);
    if (pc >= KC) return;
    if (active != nullptr && active[wg * 4 + t] == 0) return;
    double2* dst = reinterpret_cast<double2*>(
        out + (((size_t)wg * ne + ioff + il) * KC + pc) * 32 + t * 8 + (g & 3) * 2);
    *dst = make_double2(c0, c1);
  }
};

template <int WM, int WN, int CWM, int CWN, class Epi>
__global__ void __launch_bounds__(CWM* CWN * 32)
    gemm_frag_kernel(GemmArgs a, Epi epi) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int wm = warp % CWM, wn = warp / CWM;
  const int z = blockIdx.z;
  const int mt0 = (blockIdx.x * CWM + wm) * WM;
  const int nt0 = (blockIdx.y * CWN + wn) * WN;
  if (mt0 >= a.MTiles || nt0 >= a.NTiles) return;

  const double* Ap[WM];
  const double* Bp[WN];
  const int boff = b_lane_offset(lane);
#pragma unroll
  for (int i = 0; i < WM; ++i) {
    int mt = min(mt0 + i, a.MTiles - 1);
    Ap[i] = a.A + (size_t)z * a.strideAz + (size_t)mt * a.KS * 32 + lane;
  }
#pragma unroll
  for (int j = 0; j < WN; ++j) {
    int nt = min(nt0 + j, a.NTiles - 1);
    Bp[j] = a.B + (size_t)z * a.strideBz + (size_t)(nt / a.ntInner) * a.strideBO +
            (size_t)(nt % a.ntInner) * a.strideBI + boff;
  }

  double acc[WM][WN][2];
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

  double af[WM], bf[WN], an[WM], bn[WN];
#pragma unroll
  for (int i = 0; i < WM; ++i) af[i] = ldg_nc(Ap[i]);
#pragma unroll
  for (int j = 0; j < WN; ++j) bf[j] = ldg_nc(Bp[j]);

  for (int ks = 0; ks < a.KS; ++ks) {
    const int kn = (ks + 1 < a.KS) ? ks + 1 : ks;
#pragma unroll
    for (int i = 0; i < WM; ++i) an[i] = ldg_nc(Ap[i] + (size_t)kn * 32);
#pragma unroll
    for (int j = 0; j < WN; ++j) bn[j] = ldg_nc(Bp[j] + (size_t)kn * 32);
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
#pragma unroll
    for (int i = 0; i < WM; ++i) af[i] = an[i];
#pragma unroll
    for (int j = 0; j < WN; ++j) bf[j] = bn[j];
  }

  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int i = 0; i < WM; ++i) {
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      if (mt0 + i < a.MTiles && nt0 + j < a.NTiles)
        epi(mt0 + i, nt0 + j, z, g, t, acc[i][j][0], acc[i][j][1]);
    }
  }
}

template <int WM, int WN, int CWM, int CWN, class Epi>
inline cudaError_t launch_gemm(const GemmArgs& a, const Epi& epi, int batch, cudaStream_t st) {
  dim3 grid((a.MTiles + WM * CWM - 1) / (WM * CWM), (a.NTiles + WN * CWN - 1) / (WN * CWN), batch);
  // grid.y is limited to 65535: fold if needed by swapping roles is not required for
  // the shapes of this path (NTiles/ (WN*CWN) <= 65535 up to ~4M walkers x orbitals)
  gemm_frag_kernel<WM, WN, CWM, CWN, Epi><<<grid, CWM * CWN * 32, 0, st>>>(a, epi);
  return cudaGetLastError();
}

}  // namespace pxb

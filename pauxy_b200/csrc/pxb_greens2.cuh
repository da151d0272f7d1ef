// K1 (second version): overlap, inverse, determinant, rotated Green's function and one-body
// energy (walkers/single_det.py:295-321, :170-199; estimators/generic.py:178) in two kernels
//
//   1. O_s[w] = phi_s^T psi_s for ALL walkers as one fragment-major TMA GEMM (pxb_gemm.cuh,
//      epilogue EpiO): psi^T is the shared real A operand, the walker blocks of phi in their OF
//      layout are exactly its B fragments (4 walkers x (re, im) per tile).
//   2. theta_kernel: one WARP per (walker, spin), no block-level barriers:
//        - in-place Gauss-Jordan inversion with partial pivoting of the ns x ns complex O in the
//          warp's slice of shared memory (pivot order = LAPACK's izamax on |re| + |im|), slogdet
//          from the pivots;
//        - Theta = O^-1 phi^T on DMMA: A fragments from the inverse in shared memory, B fragments
//          straight from the OF layout in global memory (a fragment = four 64-byte segments), the
//          rotated partner (i B) by a lane shuffle; Theta stored back in OF layout, e1b fused.
#pragma once
#include "pxb_common.cuh"
#include "pxb_gemm.cuh"
#include "pxb_greens.cuh"

namespace pxb {

struct EpiO {  // O[(w, spin)][i][j] complex, row-major with leading dimension nld, slot stride nsq
  double2* OB;
  int ns, spin, nld, nsq;
  struct Row { int j; };           // < 0: padding rows beyond the occupied orbitals
  struct Col { double2* base; };
  __device__ __forceinline__ Row row(int mt, int g) const {
    const int j = 8 * mt + g;
    return {j < ns ? j : -1};
  }
  __device__ __forceinline__ Col col(int nt, int z, int t) const {
    const int wg = nt / ns, i = nt - wg * ns;
    return {OB + ((size_t)(4 * wg + t) * 2 + spin + z) * nsq + (size_t)i * nld};  // z: spin of a spin-batched launch
  }
  __device__ __forceinline__ void store(const Row& r, const Col& c, double c0, double c1) const {
    if (r.j >= 0) c.base[r.j] = make_double2(c0, c1);
  }
};

struct ThetaArgs {
  const double2* OB;     // overlap matrices [(w, s)][nsq]
  const double* phi;     // OF
  double* theta;         // OF
  const double2* h1rot;  // [ne][Mp]
  double* slog;          // [Wp][2][4]: sign_re, sign_im, logdet, unused
  double2* e1b_part;     // [Wp][2]
  Dims d;
  int nld, nsq;          // leading dimension (complex) and slot size of OB and of the shared copy
};

constexpr int TH_WARPS = 4;

inline size_t theta_smem_per_warp(int nmax) {
  const int nld = nmax | 1;
  return ((size_t)nmax * nld * sizeof(cplx) + (size_t)nmax * sizeof(cplx) + (size_t)nmax * sizeof(int) + 15) / 16 * 16;
}

// NMT: 8-row tiles over the occupied orbitals of one spin (ceil(ns/8) <= NMT)
template <int NMT>
__global__ void __launch_bounds__(TH_WARPS * 32, NMT <= 3 ? 7 : 1) theta_kernel(ThetaArgs a, int smem_per_warp) {
  extern __shared__ __align__(16) unsigned char th_raw[];
  const Dims& d = a.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * TH_WARPS + warp;
  const int w = item >> 1, s = item & 1;
  if (w >= d.Wp) return;
  const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
  double* sl = a.slog + ((size_t)w * 2 + s) * 4;
  if (ns == 0) {
    if (lane == 0) {
      sl[0] = 1.0;
      sl[1] = 0.0;
      sl[2] = 0.0;
      a.e1b_part[(size_t)w * 2 + s] = make_double2(0.0, 0.0);
    }
    return;
  }
  const int nmax = max(d.na, d.nb);
  const int lda = a.nld;
  cplx* A = reinterpret_cast<cplx*>(th_raw + (size_t)warp * smem_per_warp);  // [ns][lda]
  cplx* colk = A + (size_t)nmax * lda;                                        // [ns]
  int* piv = reinterpret_cast<int*>(colk + nmax);                             // [ns]
  const int g = lane >> 2, t = lane & 3;
  const int wg = w >> 2, wl = w & 3;

  // 1. O -> shared
  {
    const double2* src = a.OB + ((size_t)w * 2 + s) * a.nsq;
    for (int idx = lane; idx < ns * lda; idx += 32) {
      const double2 v = src[idx];
      A[idx] = {v.x, v.y};
    }
  }
  __syncwarp();

  // 2. in-place Gauss-Jordan inversion with partial pivoting; lanes own columns j = lane, lane + 32
  cplx sign = {1.0, 0.0};
  double logdet = 0.0;
  for (int k = 0; k < ns; ++k) {
    double bv = -1.0;
    int bi = k;
    for (int i = k + lane; i < ns; i += 32) {
      const double v = cabs1(A[(size_t)i * lda + k]);
      if (v > bv) {
        bv = v;
        bi = i;
      }
    }
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
      const double ov = __shfl_xor_sync(0xffffffffu, bv, m);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, m);
      if (ov > bv || (ov == bv && oi < bi)) {
        bv = ov;
        bi = oi;
      }
    }
    if (lane == 0) piv[k] = bi;
    if (bi != k) {
      for (int j = lane; j < ns; j += 32) {
        const cplx x = A[(size_t)k * lda + j];
        A[(size_t)k * lda + j] = A[(size_t)bi * lda + j];
        A[(size_t)bi * lda + j] = x;
      }
      sign = {-sign.re, -sign.im};
    }
    __syncwarp();
    const cplx pv = A[(size_t)k * lda + k];
    {
      const double au = hypot(pv.re, pv.im);
      sign = cmul(sign, {pv.re / au, pv.im / au});
      logdet += log(au);
    }
    const cplx rp = cdiv({1.0, 0.0}, pv);
    // column k (multipliers) aside, then row k scaled with the unit entry in place of the pivot
    for (int i = lane; i < ns; i += 32) colk[i] = A[(size_t)i * lda + k];
    __syncwarp();
    constexpr int NU = NMT > 4 ? 2 : 1;  // columns per lane: ns <= 8 NMT
    cplx rk[NU];
#pragma unroll
    for (int u = 0; u < NU; ++u) {
      const int j = lane + 32 * u;
      rk[u] = {0.0, 0.0};
      if (j < ns) {
        const cplx v = (j == k) ? cplx{1.0, 0.0} : A[(size_t)k * lda + j];
        rk[u] = cmul(v, rp);
        A[(size_t)k * lda + j] = rk[u];
      }
    }
    if (NU == 1) {
      // one column per lane: the lane predicate and the pivot-column test leave the row loop,
      // which is unrolled so that several independent shared-memory round trips are in flight
      if (lane < ns) {
        cplx* aj = A + lane;
        const cplx r0 = rk[0];
        const bool isk = lane == k;
#pragma unroll 4
        for (int i = 0; i < ns; ++i) {
          if (i == k) continue;
          const cplx f = colk[i];
          cplx v = aj[(size_t)i * lda];
          if (isk) v = {0.0, 0.0};
          aj[(size_t)i * lda] = csub(v, cmul(f, r0));
        }
      }
    } else {
      for (int i = 0; i < ns; ++i) {
        if (i == k) continue;
        const cplx f = colk[i];
#pragma unroll
        for (int u = 0; u < NU; ++u) {
          const int j = lane + 32 * u;
          if (j < ns) {
            const cplx v = (j == k) ? cplx{0.0, 0.0} : A[(size_t)i * lda + j];
            A[(size_t)i * lda + j] = csub(v, cmul(f, rk[u]));
          }
        }
      }
    }
    __syncwarp();
  }
  // undo the row interchanges as column interchanges, last first
  for (int k = ns - 1; k >= 0; --k) {
    const int p = piv[k];
    if (p != k) {
      for (int i = lane; i < ns; i += 32) {
        const cplx x = A[(size_t)i * lda + k];
        A[(size_t)i * lda + k] = A[(size_t)i * lda + p];
        A[(size_t)i * lda + p] = x;
      }
      __syncwarp();
    }
  }
  if (lane == 0) {
    sl[0] = sign.re;
    sl[1] = sign.im;
    sl[2] = logdet;
  }
  __syncwarp();

  // 3. Theta[a][p] = sum_i Oinv[a][i] phi[p][i]; two basis chunks (n-tiles) per iteration
  const int KS = (ns + 3) >> 2;
  const int nmt = (ns + 7) >> 3;
  const double* phis = a.phi + ((size_t)wg * d.ne + ioff) * d.KC * 32 + wl * 8 + g;
  const unsigned smask = (g & 1) ? 0u : 0x80000000u;  // (i B)^: (re, im) -> (-im, re)
  double er = 0.0, ei = 0.0;
  // B fragments of one iteration (two basis chunks x all k-steps) are loaded together: one
  // global-load latency per iteration instead of one per k-step.  For the larger shapes, where few
  // warps fit on an SM, they are also loaded one iteration ahead (costs 4 NMT more registers).
  constexpr int KSM = 2 * NMT;  // >= ceil(ns / 4)
  constexpr bool AHEAD = NMT >= 4;
  const double* bptr[KSM];
#pragma unroll
  for (int ks = 0; ks < KSM; ++ks) {
    const int i = min(4 * ks + t, ns - 1);  // clamped: the matching A entries are zero
    bptr[ks] = phis + (size_t)i * d.KC * 32;
  }
  double bn[KSM][2];
  auto load_b = [&](int pc0) {
    const bool one = pc0 < d.KC, two = pc0 + 1 < d.KC;
#pragma unroll
    for (int ks = 0; ks < KSM; ++ks) {
      bn[ks][0] = (one && ks < KS) ? ldg_nc(bptr[ks] + (size_t)pc0 * 32) : 0.0;
      bn[ks][1] = (two && ks < KS) ? ldg_nc(bptr[ks] + (size_t)pc0 * 32 + 32) : 0.0;
    }
  };
  if (AHEAD) load_b(0);
  for (int pc0 = 0; pc0 < d.KC; pc0 += 2) {
    if (!AHEAD) load_b(pc0);
    double b[KSM][2];
#pragma unroll
    for (int ks = 0; ks < KSM; ++ks) {
      b[ks][0] = bn[ks][0];
      b[ks][1] = bn[ks][1];
    }
    if (AHEAD) load_b(pc0 + 2);
    double acc[NMT][2][2];
#pragma unroll
    for (int m = 0; m < NMT; ++m)
#pragma unroll
      for (int q = 0; q < 2; ++q) acc[m][q][0] = acc[m][q][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < KSM; ++ks) {
      if (ks < KS) {
        double bq[2];
#pragma unroll
        for (int q = 0; q < 2; ++q) {
          const double o = __shfl_xor_sync(0xffffffffu, b[ks][q], 4);
          bq[q] = __hiloint2double(__double2hiint(o) ^ (int)smask, __double2loint(o));
        }
        const bool iv = 4 * ks + t < ns;
#pragma unroll
        for (int m = 0; m < NMT; ++m) {
          if (m < nmt) {
            const int ar = 8 * m + g;
            const cplx av = (iv && ar < ns) ? A[(size_t)ar * lda + 4 * ks + t] : cplx{0.0, 0.0};
#pragma unroll
            for (int q = 0; q < 2; ++q) {
              dmma(acc[m][q][0], acc[m][q][1], av.re, b[ks][q]);
              dmma(acc[m][q][0], acc[m][q][1], av.im, bq[q]);
            }
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < 2; ++q) {
      const int pc = pc0 + q;
      if (pc >= d.KC) continue;
      const int p = 4 * pc + t;
#pragma unroll
      for (int m = 0; m < NMT; ++m) {
        const int ar = 8 * m + g;
        if (m < nmt && ar < ns) {
          double vr = acc[m][q][0], vi = acc[m][q][1];
          if (p >= d.M) vr = vi = 0.0;
          *reinterpret_cast<double2*>(a.theta + (((size_t)wg * d.ne + ioff + ar) * d.KC + pc) * 32 + wl * 8 + t * 2) =
              make_double2(vr, vi);
          const double2 h = a.h1rot[(size_t)(ioff + ar) * d.Mp + p];
          er += h.x * vr - h.y * vi;
          ei += h.x * vi + h.y * vr;
        }
      }
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    er += __shfl_xor_sync(0xffffffffu, er, m);
    ei += __shfl_xor_sync(0xffffffffu, ei, m);
  }
  if (lane == 0) a.e1b_part[(size_t)w * 2 + s] = make_double2(er, ei);
}

}  // namespace pxb

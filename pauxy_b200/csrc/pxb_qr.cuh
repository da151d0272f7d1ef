// K8 (second version): re-orthogonalisation phi = Q R with R_ii > 0 (walkers/single_det.py:215-255,
// utils/linalg.py reortho) as CholeskyQR2 on DMMA, one warp per (walker, spin), for up to 32
// occupied orbitals per spin:
//
//   twice:  G = V^H V            Gram matrix, DMMA: the 8-orbital x 4-basis-function tiles of V in
//                                 the OF layout are A fragments (V^H) and B fragments (V) at once
//           G = R^H R            warp-level Cholesky in shared memory (R upper, real positive diagonal)
//           V <- V R^-1          warp_apply_left (the Theta product of pxb_greens2.cuh) with A = R^-T
//
// The QR factorisation with a positive diagonal is unique, so Q agrees with the reference's
// Householder QR + sign fix to rounding, and log det R = sum of the log-diagonals of both passes.
// One pass loses orthogonality like u cond(V)^2; the second pass repairs it as long as the first
// Cholesky factorisation went through with a margin (cond(V) below ~1e5, checked on the pivots).  A
// (walker, spin) that fails the check is left to the modified Gram-Schmidt kernel (qr_kernel with a
// mask), which starts from whatever this kernel left in place: V (first pass failed) or V R_1^-1
// (second pass failed; log det R_1 is then already in logdet).
#pragma once
#include "pxb_common.cuh"
#include "pxb_greens2.cuh"

namespace pxb {

struct CholQrArgs {
  double* phi;     // OF, in place
  double* logdet;  // [Wp][2]: sum_k log R_kk of this spin
  int* need_mgs;   // [Wp][2]: 1 = finish this (walker, spin) with Gram-Schmidt
  Dims d;
  int lda;         // leading dimension (complex) of the two shared matrices, max(na, nb) | 1
};

constexpr double CQR_TOL1 = 1e-10;  // first pass: pivot^2 / diagonal of G below this -> Gram-Schmidt
constexpr double CQR_TOL2 = 0.25;   // second pass: G is the identity to ~1e-5 unless the first pass was inaccurate
                                    // (both relative to the diagonal entry of G before the factorisation)

inline size_t cholqr_smem_per_warp(int nmax) {
  const int lda = nmax | 1;
  const int nc = (nmax + 7) / 8 * 8;
  // G / R [nmax][lda], A = R^-T [nmax][lda] + 3 elements of slack (apply step), + nc elements of slack: the
  // rotating-slot Cholesky reads up to nc - 1 elements past the end of a row of R (padding slots)
  return ((size_t)(2 * nmax * lda + 3 + nc) * sizeof(cplx) + 15) / 16 * 16;
}

// G = V^H V for the ns orbitals of one (walker, spin) -> Gs (full Hermitian matrix)
template <int NMT>
__device__ __forceinline__ void warp_gram(const Dims& d, const double* in, int ns, cplx* Gs, int lda, int lane) {
  const int g = lane >> 2, t = lane & 3;
  const unsigned rowB = (unsigned)d.KC * 32u;
  constexpr int NP = NMT * (NMT + 1) / 2;
  double are[NP][2], aim[NP][2];
#pragma unroll
  for (int q = 0; q < NP; ++q) are[q][0] = are[q][1] = aim[q][0] = aim[q][1] = 0.0;
  const double* vp[NMT];
  bool ok[NMT];
#pragma unroll
  for (int m = 0; m < NMT; ++m) {
    ok[m] = 8 * m + g < ns;
    vp[m] = in + (size_t)min(8 * m + g, ns - 1) * rowB + 2 * t;
  }
  double2 vn[NMT];
  auto load_v = [&](int pc) {
#pragma unroll
    for (int m = 0; m < NMT; ++m)
      vn[m] = ok[m] ? *reinterpret_cast<const double2*>(vp[m] + (size_t)pc * 32) : make_double2(0.0, 0.0);
  };
  load_v(0);
  for (int pc = 0; pc < d.KC; ++pc) {
    double2 v[NMT];
    const bool pz = 4 * pc + t >= d.M;  // basis padding of the last chunk
#pragma unroll
    for (int m = 0; m < NMT; ++m) v[m] = pz ? make_double2(0.0, 0.0) : vn[m];
    if (pc + 1 < d.KC) load_v(pc + 1);
    int q = 0;
#pragma unroll
    for (int mi = 0; mi < NMT; ++mi) {
#pragma unroll
      for (int nj = mi; nj < NMT; ++nj, ++q) {
        // lane (g, t) holds V[p = 4 pc + t][orbital 8 m + g]: A[m = g][k = t] of V^T and B[k = t][n = g] of V
        dmma(are[q][0], are[q][1], v[mi].x, v[nj].x);
        dmma(aim[q][0], aim[q][1], v[mi].x, v[nj].y);
        dmma(are[q][0], are[q][1], v[mi].y, v[nj].y);
        dmma(aim[q][0], aim[q][1], -v[mi].y, v[nj].x);
      }
    }
  }
  // C fragment: lane (g, t) holds C[g][2 t], C[g][2 t + 1]
  int q = 0;
#pragma unroll
  for (int mi = 0; mi < NMT; ++mi) {
#pragma unroll
    for (int nj = mi; nj < NMT; ++nj, ++q) {
      const int i = 8 * mi + g;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        const int j = 8 * nj + 2 * t + c;
        if (i < ns && j < ns) {
          Gs[(size_t)i * lda + j] = {are[q][c], aim[q][c]};
          if (mi != nj) Gs[(size_t)j * lda + i] = {are[q][c], -aim[q][c]};
        }
      }
    }
  }
  __syncwarp();
}

// Cholesky G = R^H R and A = R^-T for the ns x ns Hermitian matrix in Gs (ns <= NC <= 32), one warp.
// Lane j keeps column j of the upper triangle in registers; row k of R goes to shared memory (Gs is overwritten by R) so that the multipliers of
// the trailing update and the coefficients of the back substitution are broadcast loads:
//   step k:  R[k][j] = G[k][j] / sqrt(G[k][k]);  G[i][j] -= conj(R[k][i]) R[k][j]  (k < i <= j)
//   then lane j solves R x = e_j backwards and stores x as row j of A (lower triangular).
// Returns false (warp-uniform) when a pivot falls below tol times the original diagonal entry;
// ld += sum log R_kk otherwise.
template <int NC>
__device__ __forceinline__ bool warp_cholesky_inverse(cplx* Gs, cplx* As, int lda, int ns, int lane, double tol,
                                                      double& ld) {
  constexpr unsigned FULL = 0xffffffffu;
  double2 c[NC];
  const bool mine = lane < ns;
#pragma unroll
  for (int i = 0; i < NC; ++i)
    c[i] = (mine && i <= lane && i < ns) ? *reinterpret_cast<const double2*>(Gs + (size_t)i * lda + lane)
                                         : make_double2(0.0, 0.0);
  const double diag0 = mine ? Gs[(size_t)lane * lda + lane].re : 1.0;
  __syncwarp();
  double myrd = 1.0;  // 1 / R[lane][lane]
  bool bad = false;   // a failed pivot check is sticky; the arithmetic runs on (no divergence) and is discarded
  // The step loop is a runtime loop with static register indices: slot i holds row k + i at step k
  // (every update writes one slot down, as in gj_invert_regs).  Slots of rows >= ns pick up whatever
  // lies behind row k in shared memory; they never reach slot 0 and feed nothing else.
#pragma unroll 1
  for (int k = 0; k < ns; ++k) {
    const double d2 = __shfl_sync(FULL, c[0].x, k);
    bad = bad || !(d2 > tol * __shfl_sync(FULL, diag0, k));
    const double rk = rsqrt(d2);
    const double2 ck = lane == k ? make_double2(d2 * rk, 0.0) : make_double2(c[0].x * rk, c[0].y * rk);
    if (lane == k) myrd = rk;
    if (mine) *reinterpret_cast<double2*>(Gs + (size_t)k * lda + lane) = ck;  // row k of R (zero left of the diagonal)
    __syncwarp();
    const double2* rrow = reinterpret_cast<const double2*>(Gs + (size_t)k * lda + k);
#pragma unroll
    for (int i = 1; i < NC; ++i) {
      const double2 ri = rrow[i];  // R[k][k + i]
      c[i - 1].x = fma(-ri.y, ck.y, fma(-ri.x, ck.x, c[i].x));  // G[k+i][j] -= conj(R[k][k+i]) R[k][j]
      c[i - 1].y = fma(ri.y, ck.x, fma(-ri.x, ck.y, c[i].y));
    }
    c[NC - 1] = make_double2(0.0, 0.0);
    __syncwarp();  // the loads above may run past row k (padding slots): order them before the next row's store
  }
  if (bad) return false;
  {  // log det R = -sum_j log(1 / R_jj)
    double l = mine ? -log(myrd) : 0.0;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) l += __shfl_xor_sync(FULL, l, m);
    ld += l;
  }
  double2 x[NC];
#pragma unroll
  for (int i = 0; i < NC; ++i) x[i] = make_double2(i == lane ? myrd : 0.0, 0.0);
#pragma unroll
  for (int i = NC - 2; i >= 0; --i) {
    if (i < ns - 1) {
      double sr = 0.0, si = 0.0;
#pragma unroll
      for (int k = i + 1; k < NC; ++k) {
        if (k < ns) {
          const double2 r = *reinterpret_cast<const double2*>(Gs + (size_t)i * lda + k);
          sr = fma(-r.y, x[k].y, fma(r.x, x[k].x, sr));
          si = fma(r.y, x[k].x, fma(r.x, x[k].y, si));
        }
      }
      const double rd = __shfl_sync(FULL, myrd, i);
      if (lane > i) x[i] = make_double2(-sr * rd, -si * rd);
    }
  }
  // row j of A = column j of R^-1 (zero above the diagonal), plus the zeroed slack behind the last row
  if (mine) {
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (i < lda) *reinterpret_cast<double2*>(As + (size_t)lane * lda + i) = i < ns ? x[i] : make_double2(0.0, 0.0);
  }
  if (lane < 3) As[(size_t)ns * lda + lane] = {0.0, 0.0};
  __syncwarp();
  return true;
}

template <int NMT>
__global__ void __launch_bounds__(TH_WARPS * 32, NMT == 1 ? 6 : NMT == 2 ? 4 : NMT == 3 ? 3 : 2)
    cholqr_kernel(CholQrArgs a, int smem_per_warp) {
  extern __shared__ __align__(16) unsigned char cq_raw[];
  const Dims& d = a.d;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int item = blockIdx.x * TH_WARPS + warp;
  const int w = item >> 1, s = item & 1;
  if (w >= d.Wp) return;
  const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
  if (ns == 0) {
    if (lane == 0) {
      a.logdet[item] = 0.0;
      a.need_mgs[item] = 0;
    }
    return;
  }
  const int nmax = max(d.na, d.nb), lda = a.lda;
  cplx* Gs = reinterpret_cast<cplx*>(cq_raw + (size_t)warp * smem_per_warp);
  cplx* As = Gs + (size_t)nmax * lda;
  const int wg = w >> 2, wl = w & 3;
  double* V = a.phi + ((size_t)wg * d.ne + ioff) * d.KC * 32 + wl * 8;
  double ld = 0.0, er = 0.0, ei = 0.0;
  bool good = true;
#pragma unroll 1
  for (int pass = 0; pass < 2 && good; ++pass) {
    warp_gram<NMT>(d, V, ns, Gs, lda, lane);
    good = warp_cholesky_inverse<8 * NMT>(Gs, As, lda, ns, lane, pass == 0 ? CQR_TOL1 : CQR_TOL2, ld);
    if (good) {
      warp_apply_left<NMT, false, false, true>(As, lda, ns, d, V, V, nullptr, lane, er, ei);
      __syncwarp();  // the stores of this pass are visible to the loads of the next one
    }
  }
  if (lane == 0) {
    a.logdet[item] = ld;
    a.need_mgs[item] = good ? 0 : 1;
  }
}

}  // namespace pxb

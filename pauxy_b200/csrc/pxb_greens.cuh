// K1: overlap matrix, its inverse and determinant, rotated Green's function
//     Theta = O^-1 phi^T and the one-body energy, per (walker, spin) CTA.
//     walkers/single_det.py:295-321 (greens_function), :170-199 (calc_overlap),
//     estimators/generic.py:178 (e1b through the half-rotated H1).
// K8: re-orthogonalisation (walkers/single_det.py:215-255), per (walker, spin) CTA.
//
// The two small GEMMs of K1 run on DMMA out of shared memory:
//   O[i][j]      = sum_p phi[p,i] psi[p,j]            rows i, cols j, k = p   (psi real)
//   Theta[a][p]  = sum_i Oinv[a][i] phi[p,i]          rows a, cols (p, re/im), k = i
// the inverse and slogdet come from a Gauss-Jordan elimination with partial
// pivoting on the augmented matrix [O | I] (same pivots as LAPACK's LU).
#pragma once
#include "pxb_common.cuh"

namespace pxb {

struct GreensArgs {
  const double* phi;     // OF
  double* theta;         // OF
  const double* PF;      // psi B-fragments [s][JT_s][KC][32]: lane (g,t) <- psi[4pc+t][ioff + 8jt+g]
  const double2* h1rot;  // [ne][Mp]
  double* slog;          // [Wp][2][4]: sign_re, sign_im, logdet, unused
  double2* e1b_part;     // [Wp][2]
  Dims d;
  int want_theta;
};

constexpr int GR_THREADS = 128;

__device__ __forceinline__ double cabs1(cplx z) { return fabs(z.re) + fabs(z.im); }

inline __host__ __device__ int greens_ld(const Dims& d) { return d.Mp | 1; }  // odd: conflict-free rows

inline size_t greens_smem_bytes(const Dims& d) {
  const int nmax = d.na > d.nb ? d.na : d.nb;
  return sizeof(cplx) * ((size_t)nmax * greens_ld(d) + (size_t)2 * nmax * (nmax | 1) + 2 * nmax + 8) +
         sizeof(double) * 2 * (GR_THREADS / 32) + 64;
}

// NMT: 8-row tiles over the occupied orbitals of one spin (ceil(ns/8) <= NMT)
template <int NMT>
__global__ void __launch_bounds__(GR_THREADS) greens_kernel(GreensArgs a) {
  extern __shared__ __align__(16) unsigned char gs_raw[];
  const Dims& d = a.d;
  const int w = blockIdx.x >> 1, s = blockIdx.x & 1;
  const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  constexpr int NW = GR_THREADS / 32;
  if (ns == 0) {
    if (tid == 0) {
      double* sl = a.slog + ((size_t)w * 2 + s) * 4;
      sl[0] = 1.0;
      sl[1] = 0.0;
      sl[2] = 0.0;
      a.e1b_part[(size_t)w * 2 + s] = make_double2(0.0, 0.0);
    }
    return;
  }
  const int nmax = max(d.na, d.nb);
  const int LD = greens_ld(d);
  const int LDC = nmax | 1;                      // odd column stride: conflict-free column owners
  cplx* ph = reinterpret_cast<cplx*>(gs_raw);    // [ns][LD]        phi^T of this spin
  cplx* aug = ph + (size_t)nmax * LD;            // [2 ns][LDC]     COLUMN-major [O | I] -> [I | O^-1]
  cplx* colk = aug + (size_t)2 * nmax * LDC;     // [nmax]          pivot column (after the row swap)
  cplx* pvt = colk + nmax;                       // [nmax]          pivots
  cplx* bc = pvt + nmax;                         // [1]             reciprocal pivot
  double* red = reinterpret_cast<double*>(bc + 8);            // [2*NW]
  int* ipiv = reinterpret_cast<int*>(red + 2 * NW);           // [2]: pivot row, swap count
  const int wg = w >> 2, wl = w & 3;
  const int nmt = (ns + 7) >> 3;
  const cplx czero = {0.0, 0.0};

  // 1. phi_s -> shared (orbital-major); aug = [0 | I]
  for (int i = warp; i < ns; i += NW) {
    const double* src = a.phi + ((size_t)wg * d.ne + ioff + i) * d.KC * 32 + wl * 8;
    for (int p = lane; p < d.Mp; p += 32) {
      const double2 v = *reinterpret_cast<const double2*>(src + (p >> 2) * 32 + (p & 3) * 2);
      ph[i * LD + p] = {v.x, v.y};
    }
  }
  for (int j = warp; j < 2 * ns; j += NW)
    for (int i = lane; i < ns; i += 32) aug[j * LDC + i] = (j == ns + i) ? cplx{1.0, 0.0} : czero;
  if (tid == 0) ipiv[1] = 0;
  __syncthreads();

  // 2. O = phi_s^T psi_s on DMMA: tile (mt, jt), real and imaginary accumulators
  {
    const double* PFs = a.PF + (s ? (size_t)((d.na + 7) >> 3) * d.KC * 32 : 0);
    for (int tile = warp; tile < nmt * nmt; tile += NW) {
      const int mt = tile / nmt, jt = tile % nmt;
      double cr0 = 0, cr1 = 0, ci0 = 0, ci1 = 0;
      const int ia = 8 * mt + g;
      const cplx* ap = ph + (size_t)min(ia, ns - 1) * LD + t;
      const double keep = ia < ns ? 1.0 : 0.0;
      const double* bp = PFs + (size_t)jt * d.KC * 32 + lane;
#pragma unroll 3
      for (int pc = 0; pc < d.KC; ++pc) {
        const cplx av = ap[4 * pc];
        const double b = ldg_nc(bp + pc * 32);
        dmma(cr0, cr1, keep * av.re, b);
        dmma(ci0, ci1, keep * av.im, b);
      }
      const int j = 8 * jt + 2 * t;
      if (ia < ns) {
        if (j < ns) aug[j * LDC + ia] = {cr0, ci0};
        if (j + 1 < ns) aug[(j + 1) * LDC + ia] = {cr1, ci1};
      }
    }
  }
  __syncthreads();

  // 3. Gauss-Jordan with partial pivoting on [O | I]; thread j owns column j
  for (int k = 0; k < ns; ++k) {
    if (warp == 0) {
      const cplx* ck = aug + (size_t)k * LDC;
      double bv = -1.0;
      int bi = k;
      for (int i = k + lane; i < ns; i += 32) {
        const double v = cabs1(ck[i]);
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
      // column k after swapping rows k <-> bi; entry k itself is handled by the owner
      for (int i = lane; i < ns; i += 32) colk[i] = (i == k) ? czero : ck[(i == bi) ? k : i];
      if (lane == 0) {
        const cplx pv = ck[bi];
        ipiv[0] = bi;
        if (bi != k) ipiv[1] += 1;
        pvt[k] = pv;
        bc[0] = cdiv({1.0, 0.0}, pv);
      }
    }
    __syncthreads();
    const int pr = ipiv[0];
    const cplx rp = bc[0];
    for (int j = tid; j < 2 * ns; j += GR_THREADS) {
      if (j <= k) continue;  // columns 0..k of the left block are unit vectors by now: never read again
      cplx* col = aug + (size_t)j * LDC;
      const cplx xk = col[k];
      const cplx rk = cmul(col[pr], rp);   // new row k, column j
      col[pr] = xk;                        // old row k moves to row pr (no-op if pr == k)
      for (int i = 0; i < ns; ++i) col[i] = csub(col[i], cmul(colk[i], rk));
      col[k] = rk;
    }
    __syncthreads();
  }

  // 4. slogdet: sign = prod pivot/|pivot| * (-1)^swaps, logdet = sum log|pivot|
  if (tid == 0) {
    cplx sign = {(ipiv[1] & 1) ? -1.0 : 1.0, 0.0};
    double logdet = 0.0;
    for (int k = 0; k < ns; ++k) {
      const cplx u = pvt[k];
      const double au = hypot(u.re, u.im);
      sign = cmul(sign, {u.re / au, u.im / au});
      logdet += log(au);
    }
    double* sl = a.slog + ((size_t)w * 2 + s) * 4;
    sl[0] = sign.re;
    sl[1] = sign.im;
    sl[2] = logdet;
  }
  if (!a.want_theta) return;

  // 5. Theta = O^-1 phi^T on DMMA (complex x complex as two real streams), e1b partial
  double er = 0.0, ei = 0.0;
  {
    const int KS = (ns + 3) >> 2;
    const cplx* inv = aug + (size_t)ns * LDC;  // O^-1[a][i] = inv[i * LDC + a]
    for (int pc = warp; pc < d.KC; pc += NW) {
      double acc[NMT][2];
#pragma unroll
      for (int m = 0; m < NMT; ++m) acc[m][0] = acc[m][1] = 0.0;
      for (int ks = 0; ks < KS; ++ks) {
        const int i = 4 * ks + t;
        const bool iv = i < ns;
        const cplx bv = iv ? ph[(size_t)i * LD + 4 * pc + (g >> 1)] : czero;
        const double b = (g & 1) ? bv.im : bv.re;
        const double bq = (g & 1) ? bv.re : -bv.im;
#pragma unroll
        for (int m = 0; m < NMT; ++m) {
          if (m < nmt) {
            const int ar = 8 * m + g;
            const cplx av = (iv && ar < ns) ? inv[(size_t)i * LDC + ar] : czero;
            dmma(acc[m][0], acc[m][1], av.re, b);
            dmma(acc[m][0], acc[m][1], av.im, bq);
          }
        }
      }
      const int p = 4 * pc + t;
#pragma unroll
      for (int m = 0; m < NMT; ++m) {
        const int ar = 8 * m + g;
        if (m < nmt && ar < ns) {
          double vr = acc[m][0], vi = acc[m][1];
          if (p >= d.M) vr = vi = 0.0;
          *reinterpret_cast<double2*>(a.theta + (((size_t)wg * d.ne + ioff + ar) * d.KC + pc) * 32 + wl * 8 +
                                      t * 2) = make_double2(vr, vi);
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
  if (lane == 0) {
    red[warp] = er;
    red[NW + warp] = ei;
  }
  __syncthreads();
  if (tid == 0) {
    double sr = 0.0, si = 0.0;
    for (int k = 0; k < NW; ++k) {
      sr += red[k];
      si += red[NW + k];
    }
    a.e1b_part[(size_t)w * 2 + s] = make_double2(sr, si);
  }
}

// ovlp = sign_a sign_b exp(logdet_a + logdet_b) (single_det.py:321), e1b = sum of spin parts
__global__ void greens_combine_kernel(const double* __restrict__ slog, const double2* __restrict__ e1b_part,
                                      double2* __restrict__ ovlp, double2* __restrict__ e1b, int n) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const double* a = slog + (size_t)w * 8;
  const cplx sg = cmul({a[0], a[1]}, {a[4], a[5]});
  const double e = exp(a[2] + a[6]);
  ovlp[w] = make_double2(sg.re * e, sg.im * e);
  if (e1b != nullptr) {
    const double2 x = e1b_part[2 * w], y = e1b_part[2 * w + 1];
    e1b[w] = make_double2(x.x + y.x, x.y + y.y);
  }
}

// ============================================================================
// K8: QR with R_ii > 0 by modified Gram-Schmidt, one CTA per (walker, spin)
// ============================================================================
struct QrArgs {
  double* phi;      // OF, in place
  double* logdet;   // [Wp][2] sum_k log R_kk of this spin
  Dims d;
  const int* mask = nullptr;  // when set: only the (walker, spin) items it marks, and logdet is ADDED to
};

// QT threads per (walker, spin): 128 where several CTAs share an SM, 512 where the walker's orbitals fill most of
// its shared memory (one CTA per SM: the columns after v_k are then orthogonalised 16 at a time)
template <int QT>
__global__ void __launch_bounds__(QT) qr_kernel(QrArgs a) {
  extern __shared__ __align__(16) unsigned char qs_raw[];
  const Dims& d = a.d;
  if (a.mask != nullptr && a.mask[blockIdx.x] == 0) return;  // CholeskyQR2 (pxb_qr.cuh) has done this one
  const int w = blockIdx.x >> 1, s = blockIdx.x & 1;
  const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  constexpr int NW = QT / 32;
  const int LD = greens_ld(d);
  cplx* ph = reinterpret_cast<cplx*>(qs_raw);                     // [ns][LD]
  double* red = reinterpret_cast<double*>(ph + (size_t)max(ns, 1) * LD);  // [NW]
  __shared__ double s_norm;
  const int wg = w >> 2, wl = w & 3;
  if (ns == 0) {
    if (tid == 0) a.logdet[(size_t)w * 2 + s] = 0.0;
    return;
  }
  for (int idx = tid; idx < ns * d.Mp; idx += QT) {
    const int p = idx % d.Mp, i = idx / d.Mp;
    double2 v = *reinterpret_cast<const double2*>(a.phi + (((size_t)wg * d.ne + ioff + i) * d.KC + (p >> 2)) * 32 +
                                                  wl * 8 + (p & 3) * 2);
    ph[i * LD + p] = {v.x, v.y};
  }
  __syncthreads();
  double logdet = 0.0;
  for (int k = 0; k < ns; ++k) {
    cplx* vk = ph + (size_t)k * LD;
    double part = 0.0;
    for (int p = tid; p < d.M; p += QT) part += vk[p].re * vk[p].re + vk[p].im * vk[p].im;
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) part += __shfl_xor_sync(0xffffffffu, part, m);
    if (lane == 0) red[warp] = part;
    __syncthreads();
    if (tid == 0) {
      double tot = 0.0;
      for (int i = 0; i < NW; ++i) tot += red[i];
      s_norm = sqrt(tot);
    }
    __syncthreads();
    const double nrm = s_norm;
    logdet += log(nrm);
    for (int p = tid; p < d.M; p += QT) {
      vk[p].re /= nrm;
      vk[p].im /= nrm;
    }
    __syncthreads();
    // orthogonalise the later columns against v_k: one warp per column
    for (int j = k + 1 + warp; j < ns; j += NW) {
      cplx* vj = ph + (size_t)j * LD;
      double rr = 0.0, ri = 0.0;  // r = v_k^H v_j
      for (int p = lane; p < d.M; p += 32) {
        rr += vk[p].re * vj[p].re + vk[p].im * vj[p].im;
        ri += vk[p].re * vj[p].im - vk[p].im * vj[p].re;
      }
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) {
        rr += __shfl_xor_sync(0xffffffffu, rr, m);
        ri += __shfl_xor_sync(0xffffffffu, ri, m);
      }
      for (int p = lane; p < d.M; p += 32) {
        vj[p].re -= rr * vk[p].re - ri * vk[p].im;
        vj[p].im -= rr * vk[p].im + ri * vk[p].re;
      }
    }
    __syncthreads();
  }
  for (int idx = tid; idx < ns * d.Mp; idx += QT) {
    const int p = idx % d.Mp, i = idx / d.Mp;
    const cplx v = ph[i * LD + p];
    *reinterpret_cast<double2*>(a.phi + (((size_t)wg * d.ne + ioff + i) * d.KC + (p >> 2)) * 32 + wl * 8 +
                                (p & 3) * 2) = make_double2(v.re, v.im);
  }
  if (tid == 0) a.logdet[(size_t)w * 2 + s] = (a.mask != nullptr ? a.logdet[(size_t)w * 2 + s] : 0.0) + logdet;
}

inline size_t qr_smem_bytes(const Dims& d) {
  const int nmax = d.na > d.nb ? d.na : d.nb;
  return sizeof(cplx) * (size_t)nmax * greens_ld(d) + sizeof(double) * (512 / 32) + 32;
}

// detR = exp(log_det - detR_shift); log_detR += log(detR); ot = ot / detR (single_det.py:245-254).
// detR_shift != 0 only with walkers.use_log_shift; ot_true then follows the un-shifted determinant.
__global__ void qr_combine_kernel(const double* __restrict__ logdet, double2* __restrict__ ot,
                                  double2* __restrict__ ot_true, double* __restrict__ detR,
                                  double* __restrict__ log_detR, double* __restrict__ weight_free,
                                  const double* __restrict__ detR_shift, int n) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= n) return;
  const double ld = logdet[2 * w] + logdet[2 * w + 1];
  const double shift = detR_shift != nullptr ? *detR_shift : 0.0;
  const double dr = exp(ld - shift);
  detR[w] = dr;
  log_detR[w] += log(dr);
  const double2 o = ot[w];
  ot[w] = make_double2(o.x / dr, o.y / dr);
  const double dt = shift != 0.0 ? exp(ld) : dr;
  const double2 q = ot_true[w];
  ot_true[w] = make_double2(q.x / dt, q.y / dt);
  // free projection (handler.py:178-181): polar(detR) with detR real and positive -> the
  // magnitude goes into the weight, the phase factor is exactly 1
  if (weight_free != nullptr) weight_free[w] *= dr;
}

}  // namespace pxb

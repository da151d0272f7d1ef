// Exchange energy  exx_s[w] = sum_x sum_ij T_s[x,i,j] T_s[x,j,i],
// T_s[x] = R_s[x] Theta_s^T   (estimators/generic.py:198-214), fused so that T
// (57.8 GB at c4) is never materialised.
//
// Tile mapping: the 8 rows of a DMMA tile are 8 Cholesky indices x for ONE
// occupied orbital i, the 8 columns are 4 walkers x (re, im) for ONE occupied
// orbital j.  Lane (g,t) of tile (i,j) therefore holds the complex number
// T[x0+g, i, j] of walker w0+t, and the same lane of tile (j,i) holds its
// transposed partner: the trace is thread-local, and occupied-orbital counts
// that are not multiples of 8 (5, 7, 21, 40) cost no padding flops.
//
// A warp work item is a pair of orbital blocks (I, J), |I|,|J| <= BS: it
// accumulates the I x J and J x I tiles (2*|I|*|J| DMMA per k-step for
// 2*(|I|+|J|) fragment loads).  Items of one (walker group, spin) unit are
// dealt round-robin to the 8 warps of a persistent CTA; Theta of the unit
// sits in shared memory (one TMA bulk copy), R fragments stream from L2.
#pragma once
#include "pxb_common.cuh"

namespace pxb {

struct ExArgs {
  const double* RF;     // half-rotated Cholesky, fragment layout (both spins)
  const double* theta;  // OF layout
  double* exx;          // complex [2][Wp]
  Dims d;
  int smem_b_doubles;   // capacity of the shared Theta buffer (0: stream from global)
};

constexpr int EX_WARPS = 8;
constexpr int EX_MAX_BLOCKS = 16;  // orbital blocks per spin (<= 64 occupied orbitals at BS = 4)

// One work item with compile-time block sizes: NI x NJ tiles of T[.., I, J] and, off
// the diagonal, NJ x NI tiles of T[.., J, I].  No predication inside the k-loop: a
// predicated-off DMMA still occupies the FP64 pipe (measured: 72 % pipe-active at 55 %
// useful rate with run-time block sizes).
template <int NI, int NJ, bool DIAG, bool BSMEM>
__device__ __forceinline__ void exchange_item(const double* __restrict__ aI,
                                              const double* __restrict__ aJ,
                                              const double* __restrict__ bI,
                                              const double* __restrict__ bJ, int rowstride, int KC,
                                              double& sum_re, double& sum_im) {
  constexpr int NJ2 = DIAG ? 1 : NJ;  // J-side fragments are unused on the diagonal
  double acc1[NI][NJ][2], acc2[NJ2][NI][2];
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc1[i][j][0] = acc1[i][j][1] = 0.0;
#pragma unroll
  for (int j = 0; j < NJ2; ++j)
#pragma unroll
    for (int i = 0; i < NI; ++i) acc2[j][i][0] = acc2[j][i][1] = 0.0;

  double fa_i[NI], fb_i[NI], fa_j[NJ2], fb_j[NJ2];
  double na_i[NI], nb_i[NI], na_j[NJ2], nb_j[NJ2];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    fa_i[i] = ldg_nc(aI + i * rowstride);
    fb_i[i] = BSMEM ? bI[i * rowstride] : ldg_nc(bI + i * rowstride);
  }
  if (!DIAG) {
#pragma unroll
    for (int j = 0; j < NJ2; ++j) {
      fa_j[j] = ldg_nc(aJ + j * rowstride);
      fb_j[j] = BSMEM ? bJ[j * rowstride] : ldg_nc(bJ + j * rowstride);
    }
  }
  for (int pc = 0; pc < KC; ++pc) {
    const int pn = (pc + 1 < KC ? pc + 1 : pc) * 32;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      na_i[i] = ldg_nc(aI + i * rowstride + pn);
      nb_i[i] = BSMEM ? bI[i * rowstride + pn] : ldg_nc(bI + i * rowstride + pn);
    }
    if (!DIAG) {
#pragma unroll
      for (int j = 0; j < NJ2; ++j) {
        na_j[j] = ldg_nc(aJ + j * rowstride + pn);
        nb_j[j] = BSMEM ? bJ[j * rowstride + pn] : ldg_nc(bJ + j * rowstride + pn);
      }
    }
#pragma unroll
    for (int i = 0; i < NI; ++i)
#pragma unroll
      for (int j = 0; j < NJ; ++j)
        dmma(acc1[i][j][0], acc1[i][j][1], fa_i[i], DIAG ? fb_i[j] : fb_j[j]);
    if (!DIAG) {
#pragma unroll
      for (int j = 0; j < NJ2; ++j)
#pragma unroll
        for (int i = 0; i < NI; ++i) dmma(acc2[j][i][0], acc2[j][i][1], fa_j[j], fb_i[i]);
    }
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      fa_i[i] = na_i[i];
      fb_i[i] = nb_i[i];
    }
    if (!DIAG) {
#pragma unroll
      for (int j = 0; j < NJ2; ++j) {
        fa_j[j] = na_j[j];
        fb_j[j] = nb_j[j];
      }
    }
  }
  // thread-local trace: lane (g,t) holds T[x0+g, i, j] of walker 4wg+t
  double pr = 0.0, pi = 0.0;
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const double ar = acc1[i][j][0], ai = acc1[i][j][1];
      const double br = DIAG ? acc1[j][i][0] : acc2[j][i][0];
      const double bi = DIAG ? acc1[j][i][1] : acc2[j][i][1];
      pr += ar * br - ai * bi;
      pi += ar * bi + ai * br;
    }
  const double f = DIAG ? 1.0 : 2.0;
  sum_re += f * pr;
  sum_im += f * pi;
}

template <int BS, bool BSMEM>
__global__ void __launch_bounds__(EX_WARPS * 32, 1) exchange_kernel(ExArgs a) {
  extern __shared__ __align__(128) unsigned char smem_raw[];
  double* Bs = reinterpret_cast<double*>(smem_raw);
  unsigned char* tail = smem_raw + (size_t)a.smem_b_doubles * 8;
  uint64_t* bar = reinterpret_cast<uint64_t*>(tail);
  cplx* red = reinterpret_cast<cplx*>(tail + 16);               // [EX_WARPS][4]
  unsigned char* pbI = tail + 16 + EX_WARPS * 4 * sizeof(cplx);  // pair-block tables
  unsigned char* pbJ = pbI + EX_MAX_BLOCKS * (EX_MAX_BLOCKS + 1) / 2;

  const Dims& d = a.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int boff = b_lane_offset(lane);
  unsigned phase = 0;

  if (BSMEM && tid == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  __syncthreads();

  const int nunits = 2 * d.WG;
  for (int unit = blockIdx.x; unit < nunits; unit += gridDim.x) {
    const int wg = unit >> 1, s = unit & 1;
    const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
    if (ns == 0) {
      if (tid < 4) {
        double2* e = reinterpret_cast<double2*>(a.exx) + (size_t)s * d.Wp + wg * 4 + tid;
        *e = make_double2(0.0, 0.0);
      }
      continue;
    }
    const int nblk = (ns + BS - 1) / BS;
    const int base = ns / nblk, rem = ns % nblk;
    const int npb = nblk * (nblk + 1) / 2;
    __syncthreads();  // previous unit fully consumed (shared Theta, tables, red)
    if (tid == 0) {
      int k = 0;
      for (int bi = 0; bi < nblk; ++bi)
        for (int bj = bi; bj < nblk; ++bj) {
          pbI[k] = (unsigned char)bi;
          pbJ[k] = (unsigned char)bj;
          ++k;
        }
    }
    const double* Bg = a.theta + ((size_t)wg * d.ne + ioff) * d.KC * 32;
    if (BSMEM) {
      if (tid == 0) {
        fence_proxy_async();
        const unsigned total = (unsigned)((size_t)ns * d.KC * 32 * 8);
        mbar_expect_tx(bar, total);
        unsigned off = 0;
        while (off < total) {
          unsigned chunk = min(total - off, 32768u);
          tma_bulk_g2s(reinterpret_cast<unsigned char*>(Bs) + off,
                       reinterpret_cast<const unsigned char*>(Bg) + off, chunk, bar);
          off += chunk;
        }
      }
    }
    __syncthreads();  // tables visible
    if (BSMEM) {
      mbar_wait(bar, phase);
      phase ^= 1;
    }
    const double* Bsrc = BSMEM ? Bs : Bg;
    const double* RFs = a.RF + rf_spin_base(d, s);

    double sum_re = 0.0, sum_im = 0.0;
    const int nitems = d.XG * npb;
    const int rowstride = d.KC * 32;
    for (int it = warp; it < nitems; it += EX_WARPS) {
      const int xg = it / npb, pb = it % npb;
      const int bI = pbI[pb], bJ = pbJ[pb];
      const int I0 = bI * base + min(bI, rem), ni = base + (bI < rem ? 1 : 0);
      const int J0 = bJ * base + min(bJ, rem), nj = base + (bJ < rem ? 1 : 0);
      const double* aI = RFs + ((size_t)xg * ns + I0) * d.KC * 32 + lane;
      const double* aJ = RFs + ((size_t)xg * ns + J0) * d.KC * 32 + lane;
      const double* bIp = Bsrc + (size_t)I0 * d.KC * 32 + boff;
      const double* bJp = Bsrc + (size_t)J0 * d.KC * 32 + boff;
      const int code = (bI == bJ ? 16 : 0) + (ni - 1) * 4 + (nj - 1);
#define PXB_EX_CASE(NI_, NJ_)                                                                       \
  case (NI_ - 1) * 4 + (NJ_ - 1):                                                                   \
    exchange_item<NI_, NJ_, false, BSMEM>(aI, aJ, bIp, bJp, rowstride, d.KC, sum_re, sum_im);      \
    break;
#define PXB_EX_DIAG(NI_)                                                                            \
  case 16 + (NI_ - 1) * 4 + (NI_ - 1):                                                              \
    exchange_item<NI_, NI_, true, BSMEM>(aI, aJ, bIp, bJp, rowstride, d.KC, sum_re, sum_im);       \
    break;
      switch (code) {
        PXB_EX_CASE(1, 1) PXB_EX_CASE(1, 2) PXB_EX_CASE(2, 1) PXB_EX_CASE(2, 2)
        PXB_EX_CASE(2, 3) PXB_EX_CASE(3, 2) PXB_EX_CASE(3, 3) PXB_EX_CASE(3, 4)
        PXB_EX_CASE(4, 3) PXB_EX_CASE(4, 4)
        PXB_EX_DIAG(1) PXB_EX_DIAG(2) PXB_EX_DIAG(3) PXB_EX_DIAG(4)
        default: __trap();
      }
#undef PXB_EX_CASE
#undef PXB_EX_DIAG
    }
    // reduce over the 8 x-lanes (g) of each walker column t, then over warps
#pragma unroll
    for (int m = 4; m < 32; m <<= 1) {
      sum_re += __shfl_xor_sync(0xffffffffu, sum_re, m);
      sum_im += __shfl_xor_sync(0xffffffffu, sum_im, m);
    }
    if (g == 0) red[warp * 4 + t] = {sum_re, sum_im};
    __syncthreads();
    if (tid < 4) {
      double r = 0.0, i = 0.0;
      for (int w = 0; w < EX_WARPS; ++w) {
        r += red[w * 4 + tid].re;
        i += red[w * 4 + tid].im;
      }
      double2* e = reinterpret_cast<double2*>(a.exx) + (size_t)s * d.Wp + wg * 4 + tid;
      *e = make_double2(r, i);
    }
  }
}

inline size_t exchange_tail_bytes() {
  return 16 + EX_WARPS * 4 * sizeof(cplx) + 2 * (EX_MAX_BLOCKS * (EX_MAX_BLOCKS + 1) / 2) + 64;
}

}  // namespace pxb

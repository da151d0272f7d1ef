// Exchange energy as a quadratic form in the half-rotated two-electron integrals.
//
//   exx_s[w] = sum_{x,i,j} T[x,i,j] T[x,j,i],   T[x] = R_s[x] Theta_s^T      (estimators/generic.py:198-214)
//            = sum_{(i,q),(j,p)} Theta[i,q] K[(i,q),(j,p)] Theta[j,p]
//   K[(i,q),(j,p)] = sum_x R[(i,p),x] R[(j,q),x]
//
// K is the reference's half-rotated ERI v[i,p,j,q] = sum_x R[(i,p),x] R[(j,q),x] with the two
// basis indices swapped, i.e. the contraction of local_energy_generic_opt
// (estimators/generic.py:146-147: eK = -1/2 sum v[i,r,j,s] G[i,s] G[j,r]) with the ERI rebuilt
// from the SAME Cholesky vectors, so it equals the Cholesky-form exchange to rounding while
// costing 4 (ns M)^2 real flops per walker instead of 4 N ns^2 M  (N/M ~ 5x fewer), and half
// of that again because K is symmetric:
//
//   theta^T K theta = sum_A theta_A^T K_AA theta_A + 2 sum_{A<B} theta_A^T K_AB theta_B
//
// One GEMM  Y = K Theta  over the upper block triangle with the quadratic-form epilogue fused:
// Y is never written.  Same machinery as the fragment-major TMA GEMM (pxb_gemm.cuh): persistent
// CTAs, one producer warp feeding a 3-stage shared-memory ring with 1-D bulk copies, 8 consumer
// warps of 4 x 8 DMMA tiles.  The B operand is Theta in its OF layout (k-step = (orbital j,
// basis chunk pc)), exactly the operand of the force-bias GEMM.
#pragma once
#include "pxb_common.cuh"
#include "pxb_gemm.cuh"

namespace pxb {

constexpr int EQ_WM = 4, EQ_WN = 8, EQ_CWM = 4, EQ_CWN = 2;
constexpr int EQ_TM = EQ_WM * EQ_CWM;              // 16 m-tiles = 128 rows per row block
constexpr int EQ_TN = EQ_WN * EQ_CWN;              // 16 n-tiles = 64 walkers per walker block
constexpr int EQ_DIAG_STAGES = EQ_TM * 2 / GT_KS;  // ring stages covering the diagonal block
static_assert(EQ_TM * 2 % GT_KS == 0, "diagonal block must be a whole number of stages");

struct EriArgs {
  const double* KF[2];  // per spin [MT_s][KS_s][32] A-fragments of K
  const double* theta;  // OF
  double2* part;        // [2][nslot][Wp] partial sums (slot = row block * CWM + warp row)
  Dims d;
  int nslot;
};

__host__ __device__ inline int eri_dim(const Dims& d, int s) { return (s ? d.nb : d.na) * d.Mp; }
__host__ __device__ inline int eri_mtiles(const Dims& d, int s) { return (eri_dim(d, s) + 7) / 8; }
__host__ __device__ inline int eri_rowblocks(const Dims& d, int s) {
  return (eri_mtiles(d, s) + EQ_TM - 1) / EQ_TM;
}
__host__ __device__ inline size_t eri_kf_doubles(const Dims& d, int s) {
  return (size_t)eri_mtiles(d, s) * (size_t)((s ? d.nb : d.na) * d.KC) * 32;
}
inline size_t eri_smem_bytes() { return gemm_tma_smem_bytes<EQ_WM, EQ_WN, EQ_CWM, EQ_CWN>() + 2 * 4 * 8 + 4 * 4 + 16; }

// item -> (walker super-block, row block, spin, walker block inside the super-block).  A super-block
// is a run of SB consecutive walker blocks whose Theta (4.6 MB per block at c4) fits in L2 next to K
// (41-82 MB at c4): its Theta is read from DRAM once and then served by L2 to all the row blocks,
// while K stays L2-resident because every row block of it is streamed all the time.  (With the row
// block outermost over ALL walkers each walker block's Theta came back from DRAM once per row block:
// 8.4 x the algorithmic traffic.)  Inside a super-block the items are ordered longest k-range first
// (row block 0 first; lengths differ by a factor of 25) and handed out from a global counter, so the
// CTAs finish level to within one short item also when there are only a few items per CTA
// (1024 walkers per GPU: 576 items on 148 CTAs).
struct EriItem {
  int rb, s, nb;
  int mt0, nt0, ks0, nkstage, MT, KS;
  bool valid;
};
__device__ __forceinline__ EriItem eri_item(const Dims& d, int item, int nrb2, int SB) {
  EriItem it;
  const int nwb = (d.WG + EQ_TN - 1) / EQ_TN;
  const int sb = item / (nrb2 * SB);             // all super-blocks but the last hold SB walker blocks
  const int rem = item - sb * nrb2 * SB;
  const int nin = min(SB, nwb - sb * SB);
  const int r = rem / nin;
  it.nb = sb * SB + (rem - r * nin);             // walker block
  it.s = r & 1;
  it.rb = r >> 1;
  it.MT = eri_mtiles(d, it.s);
  it.KS = (it.s ? d.nb : d.na) * d.KC;
  it.mt0 = it.rb * EQ_TM;
  it.nt0 = it.nb * EQ_TN;
  it.ks0 = it.rb * EQ_TM * 2;  // k-step of the first column of the diagonal block
  it.valid = it.mt0 < it.MT;
  it.nkstage = it.valid ? (it.KS - it.ks0 + GT_KS - 1) / GT_KS : 0;
  return it;
}

constexpr int EQ_QD = 4;  // depth of the producer -> consumer item queue

static_assert(EQ_CWM * EQ_CWN == 8, "register rebalancing below assumes two consumer warpgroups");
__global__ void __launch_bounds__(gemm_tma_threads<EQ_CWM * EQ_CWN>(), 1) exx_eri_kernel(EriArgs a, int nitems, int nrb2, int SB, int* __restrict__ counter) {
  constexpr int TM = EQ_TM, TN = EQ_TN, NCW = EQ_CWM * EQ_CWN, WM = EQ_WM, WN = EQ_WN;
  constexpr int A_STAGE = TM * GT_KS * 32, B_STAGE = TN * GT_KS * 32;
  extern __shared__ __align__(128) double eq_smem[];
  double* As = eq_smem;
  double* Bs = eq_smem + GT_STAGES * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(Bs + GT_STAGES * B_STAGE);
  uint64_t* empty = full + GT_STAGES;
  uint64_t* qfull = empty + GT_STAGES;   // item queue: the producer announces the items it draws
  uint64_t* qempty = qfull + EQ_QD;
  volatile int* qitem = reinterpret_cast<volatile int*>(qempty + EQ_QD);
  const Dims& d = a.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GT_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCW * kReleaseArrivals);
    }
#pragma unroll
    for (int q = 0; q < EQ_QD; ++q) {
      mbar_init(&qfull[q], 1);
      mbar_init(&qempty[q], NCW * kReleaseArrivals);
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp >= NCW) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GT_REGS_PRODUCER));
    if (warp != NCW) return;
    // ---------------- producer: lane L streams row L of the stage (A rows, then B rows) -----------
    static_assert(TM + TN <= 32, "one producer lane per tile row");
    unsigned itc = 0, qc = 0;
    for (;;) {
      // next item from the global counter (invalid combinations of an uneven spin pair are skipped
      // here, so consumers only ever see real items); -1 ends the kernel
      int item = -1;
      for (;;) {
        if (lane == 0) item = atomicAdd(counter, 1);
        item = __shfl_sync(0xffffffffu, item, 0);
        if (item >= nitems) {
          item = -1;
          break;
        }
        if (eri_item(d, item, nrb2, SB).valid) break;
      }
      {
        const unsigned q = qc % EQ_QD, qph = (qc / EQ_QD) & 1u;
        mbar_wait(&qempty[q], qph ^ 1u);
        if (lane == 0) {
          qitem[q] = item;
          mbar_arrive(&qfull[q]);   // release: the slot's content is visible to whoever sees the phase flip
        }
        ++qc;
      }
      if (item < 0) break;
      const EriItem it = eri_item(d, item, nrb2, SB);
      const int ioff = it.s ? d.na : 0;
      const int rows_m = min(TM, it.MT - it.mt0), rows_n = min(TN, d.WG - it.nt0);
      const double* src = nullptr;
      int dst_off = 0;
      const bool isA = lane < TM;
      if (isA) {
        if (lane < rows_m) src = a.KF[it.s] + (size_t)(it.mt0 + lane) * it.KS * 32;
        dst_off = lane * GT_KS * 32;
      } else {
        const int j = lane - TM;
        if (j < rows_n) src = a.theta + ((size_t)(it.nt0 + j) * d.ne + ioff) * d.KC * 32;
        dst_off = j * GT_KS * 32;
      }
      for (int ks = 0; ks < it.nkstage; ++ks, ++itc) {
        const unsigned s = itc % GT_STAGES, ph = (itc / GT_STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        const int k0 = it.ks0 + ks * GT_KS, nk = min(GT_KS, it.KS - k0);
        const unsigned rowbytes = (unsigned)nk * 256u;
        if (lane == 0) mbar_expect_tx(&full[s], (unsigned)(rows_m + rows_n) * rowbytes);
        __syncwarp();
        if (src != nullptr) {
          double* dst = (isA ? As + (size_t)s * A_STAGE : Bs + (size_t)s * B_STAGE) + dst_off;
          tma_bulk_g2s(dst, src + (size_t)k0 * 32, rowbytes, &full[s]);
        }
      }
    }
    return;
  }
  // ---------------- consumers ----------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GT_REGS_CONSUMER));
  const int wm = warp % EQ_CWM, wn = warp / EQ_CWM;
  const int g = lane >> 2, t = lane & 3;
  const int boff = b_lane_offset(lane);
  unsigned itc = 0, qc = 0;
  for (;;) {
    int item;
    {
      const unsigned q = qc % EQ_QD, qph = (qc / EQ_QD) & 1u;
      mbar_wait(&qfull[q], qph);
      item = qitem[q];
      ring_release(&qempty[q], lane);
      ++qc;
    }
    if (item < 0) break;
    const EriItem it = eri_item(d, item, nrb2, SB);
    double acc[WM][WN][2];
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    // the Theta elements of the quadratic-form epilogue (one per accumulator) are requested now:
    // their L2 latency is hidden behind the whole k-loop instead of sitting in front of the epilogue
    const int ioff = it.s ? d.na : 0;
    const int D = eri_dim(d, it.s);
    const int mtw = it.mt0 + wm * WM, ntw = it.nt0 + wn * WN;
    double2 th[WM][WN];
    bool rowok[WM];
#pragma unroll
    for (int i = 0; i < WM; ++i) {
      const int arow = 8 * (mtw + i) + g;
      rowok[i] = (mtw + i < it.MT) && (arow < D);
      const int orb = rowok[i] ? arow / d.Mp : 0;
      const int q = rowok[i] ? arow - orb * d.Mp : 0;
      const double* trow = a.theta + ((size_t)(ioff + orb) * d.KC + (q >> 2)) * 32 + t * 8 + (q & 3) * 2;
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        const int wg = min(ntw + j, d.WG - 1);
        th[i][j] = rowok[i] ? ldg_nc2(trow + (size_t)wg * d.ne * d.KC * 32) : make_double2(0.0, 0.0);
      }
    }
    for (int ks = 0; ks < it.nkstage; ++ks, ++itc) {
      const unsigned s = itc % GT_STAGES, ph = (itc / GT_STAGES) & 1u;
      mbar_wait(&full[s], ph);
      const double* as = As + (size_t)s * A_STAGE + (size_t)wm * WM * GT_KS * 32 + lane;
      const double* bs = Bs + (size_t)s * B_STAGE + (size_t)wn * WN * GT_KS * 32 + boff;
      const int nk = min(GT_KS, it.KS - it.ks0 - ks * GT_KS);
      if (nk == GT_KS) {
#pragma unroll
        for (int kk = 0; kk < GT_KS; ++kk) {
          double af[WM], bf[WN];
#pragma unroll
          for (int i = 0; i < WM; ++i) af[i] = as[(i * GT_KS + kk) * 32];
#pragma unroll
          for (int j = 0; j < WN; ++j) bf[j] = bs[(j * GT_KS + kk) * 32];
#pragma unroll
          for (int i = 0; i < WM; ++i)
#pragma unroll
            for (int j = 0; j < WN; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      } else {
        for (int kk = 0; kk < nk; ++kk) {
          double af[WM], bf[WN];
#pragma unroll
          for (int i = 0; i < WM; ++i) af[i] = as[(i * GT_KS + kk) * 32];
#pragma unroll
          for (int j = 0; j < WN; ++j) bf[j] = bs[(j * GT_KS + kk) * 32];
#pragma unroll
          for (int i = 0; i < WM; ++i)
#pragma unroll
            for (int j = 0; j < WN; ++j) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
        }
      }
      ring_release(&empty[s], lane);
    }
    // quadratic-form epilogue: lane (g,t) of tile (i,j) holds Y[a = 8(mt0+i)+g] of walker 4(nt0+j)+t
    // the diagonal block counts once, everything to its right twice (K symmetric): KF stores the
    // diagonal blocks pre-multiplied by 1/2 (exact), so one factor serves the whole row block and
    // the accumulators are never touched between DMMAs
    const double factor = 2.0;
#pragma unroll
    for (int j = 0; j < WN; ++j) {
      const int wg = ntw + j;
      if (wg >= d.WG) continue;  // uniform over the warp
      double sr = 0.0, si = 0.0;
#pragma unroll
      for (int i = 0; i < WM; ++i) {
        if (rowok[i]) {
          sr += th[i][j].x * acc[i][j][0] - th[i][j].y * acc[i][j][1];
          si += th[i][j].x * acc[i][j][1] + th[i][j].y * acc[i][j][0];
        }
      }
#pragma unroll
      for (int m = 4; m < 32; m <<= 1) {
        sr += __shfl_xor_sync(0xffffffffu, sr, m);
        si += __shfl_xor_sync(0xffffffffu, si, m);
      }
      if (g == 0)
        a.part[((size_t)it.s * a.nslot + it.rb * EQ_CWM + wm) * d.Wp + 4 * wg + t] =
            make_double2(factor * sr, factor * si);
    }
  }
}

// exx[s][w] = sum over the slots of spin s, fixed order (deterministic)
// part_i (complex K only): the partial sums of theta^T Ki theta; exx = Er + i Ei
__global__ void exx_eri_reduce_kernel(const double2* __restrict__ part, const double2* __restrict__ part_i,
                                      double2* __restrict__ exx, Dims d, int nslot, int* __restrict__ counter) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx == 0) *counter = 0;  // re-arm the item scheduler of exx_eri_kernel for its next launch
  if (idx >= 2 * d.Wp) return;
  const int s = idx / d.Wp, w = idx % d.Wp;
  const int ns = eri_rowblocks(d, s) * EQ_CWM;
  double r = 0.0, i = 0.0;
  for (int k = 0; k < ns; ++k) {
    const double2 v = part[((size_t)s * nslot + k) * d.Wp + w];
    r += v.x;
    i += v.y;
  }
  if (part_i != nullptr) {
    double ri = 0.0, ii = 0.0;
    for (int k = 0; k < ns; ++k) {
      const double2 v = part_i[((size_t)s * nslot + k) * d.Wp + w];
      ri += v.x;
      ii += v.y;
    }
    r -= ii;
    i += ri;
  }
  exx[idx] = make_double2(r, i);
}

// ---------------------------------------------------------------------------------------------
// setup: K in A-fragment order from trial._rchol (c128 [(na+nb) M, N], real-valued).
// One CTA per (orbital pair (i,j), 32 x 32 tile of (q,p)):  K[(i,q),(j,p)] = sum_x R_i[p,x] R_j[q,x]
// ---------------------------------------------------------------------------------------------
// part 0: real Cholesky vectors; complex ones (K = Kr + i Ki, both symmetric): part 1 -> Kr, part 2 -> Ki
__global__ void __launch_bounds__(1024) eri_build_kernel(const double2* __restrict__ rchol, double* __restrict__ KF,
                                                         Dims d, int s, int part) {
  __shared__ double sq[32][33], sp[32][33], sqi[32][33], spi[32][33];
  const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
  const int tiles = (d.M + 31) / 32;
  const int qt = blockIdx.x % tiles, pt = blockIdx.x / tiles;
  const int i = blockIdx.y, j = blockIdx.z;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const int q = qt * 32 + ty, p = pt * 32 + tx;  // this thread's output element
  const int KS = ns * d.KC;
  double acc = 0.0;
  for (int x0 = 0; x0 < d.N; x0 += 32) {
    const int x = x0 + tx;
    const int qr = qt * 32 + ty, pr = pt * 32 + ty;
    const double2 zq = (qr < d.M && x < d.N) ? rchol[((size_t)(ioff + j) * d.M + qr) * d.N + x] : make_double2(0.0, 0.0);
    const double2 zp = (pr < d.M && x < d.N) ? rchol[((size_t)(ioff + i) * d.M + pr) * d.N + x] : make_double2(0.0, 0.0);
    sq[ty][tx] = zq.x;
    sp[ty][tx] = zp.x;
    if (part != 0) {
      sqi[ty][tx] = zq.y;
      spi[ty][tx] = zp.y;
    }
    __syncthreads();
    if (part == 0) {
#pragma unroll
      for (int k = 0; k < 32; ++k) acc += sq[ty][k] * sp[tx][k];
    } else if (part == 1) {
#pragma unroll
      for (int k = 0; k < 32; ++k) acc += sq[ty][k] * sp[tx][k] - sqi[ty][k] * spi[tx][k];
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k) acc += sq[ty][k] * spi[tx][k] + sqi[ty][k] * sp[tx][k];
    }
    __syncthreads();
  }
  if (q < d.M && p < d.M) {
    const int arow = i * d.Mp + q;
    const int ks = j * d.KC + (p >> 2);
    // diagonal block of the row block this row belongs to (columns [128 rb, 128 rb + 128)): x 1/2
    const int rb = arow / (EQ_TM * 8);
    const bool diag = ks >= rb * EQ_TM * 2 && ks < (rb + 1) * EQ_TM * 2;
    KF[((size_t)(arow >> 3) * KS + ks) * 32 + (arow & 7) * 4 + (p & 3)] = diag ? 0.5 * acc : acc;
  }
}

// flag[0] |= 8 unless the two spin blocks of RF are identical (then K is shared between spins)
__global__ void rf_spin_compare_kernel(const double* __restrict__ RF, size_t n, size_t base1, int* flag) {
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x)
    if (RF[idx] != RF[base1 + idx]) {
      atomicOr(flag, 8);
      return;
    }
}

}  // namespace pxb

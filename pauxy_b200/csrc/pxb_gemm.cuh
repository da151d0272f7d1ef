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

// Epilogues.  A warp owns WM x WN tiles: everything that depends only on the tile row (mt) or only
// on the tile column (nt) -- index divisions, map look-ups, base pointers -- is computed once per
// row / column (row(), col()) and store() is left with an add and a store.  (Before this split the
// one-body GEMM executed as many instructions in its epilogue as in its main loop: two integer
// divisions per tile.)
struct EpiX {  // X[z][w][n] complex128, ld = Np
  double* X;
  int Wp, Np;
  struct Row { int n; };
  struct Col { double2* base; };
  __device__ __forceinline__ Row row(int mt, int g) const { return {mt * 8 + g}; }
  __device__ __forceinline__ Col col(int nt, int z, int t) const {
    return {reinterpret_cast<double2*>(X) + ((size_t)z * Wp + nt * 4 + t) * Np};
  }
  __device__ __forceinline__ void store(const Row& r, const Col& c, double c0, double c1) const {
    c.base[r.n] = make_double2(c0, c1);
  }
};

struct EpiVHS {  // VF[w][MT][KC][c][g][t]; row tile mt = (mtv*4+s)*KC + kc
  double* VF;
  int KC, MT;
  double sqrt_dt;
  size_t walker_stride;
  const int* rt_map;  // symmetric L: compact row-tile list (upper triangle); else null
  struct Row { int off, moff; };  // element offsets inside a walker's VF block; moff < 0: no mirror
  struct Col { double* vw; };
  __device__ __forceinline__ Row row(int rt, int g) const {
    if (rt_map != nullptr) rt = rt_map[rt];
    const int kc = rt % KC, ms = rt / KC, s = ms & 3, mtv = ms >> 2;
    Row r;
    r.off = (mtv * KC + kc) * 64 + 8 * s + g;
    r.moff = -1;
    if (rt_map != nullptr) {
      // mirror VHS[q][p] = VHS[p][q] (complex symmetric, no conjugation)
      const int p = 8 * mtv + 2 * s + (g >> 2), q = 4 * kc + (g & 3);
      if (q > p && (q >> 3) < MT) r.moff = ((q >> 3) * KC + (p >> 2)) * 64 + (q & 7) * 4 + (p & 3);
    }
    return r;
  }
  __device__ __forceinline__ Col col(int nt, int z, int t) const {
    return {VF + (size_t)(nt * 4 + t) * walker_stride};
  }
  __device__ __forceinline__ void store(const Row& r, const Col& c, double c0, double c1) const {
    // VHS = i sqrt(dt) (S_re + i S_im)
    const double vr = -sqrt_dt * c1, vi = sqrt_dt * c0;
    double* base = c.vw + r.off;
    base[0] = vr;
    base[32] = vi;
    if (r.moff >= 0) {
      double* m = c.vw + r.moff;
      m[0] = vr;
      m[32] = vi;
    }
  }
};

struct EpiOF {  // out OF buffer; n-tile nt = wg*nInner + il, orbital i = ioff + il
  double* out;
  const int* active;  // optional per-walker mask (skip store if 0)
  int ne, KC, ioff, nInner;
  struct Row { int off; };        // < 0: padding rows beyond the basis
  struct Col { double* base; };   // nullptr: inactive walker, nothing stored
  __device__ __forceinline__ Row row(int mt, int g) const {
    const int pc = 2 * mt + (g >> 2);
    return {pc < KC ? pc * 32 + (g & 3) * 2 : -1};
  }
  __device__ __forceinline__ Col col(int nt, int z, int t) const {
    const int wg = nt / nInner, il = nt - wg * nInner;
    if (active != nullptr && active[wg * 4 + t] == 0) return {nullptr};
    return {out + ((size_t)wg * ne + ioff + z * nInner + il) * KC * 32 + t * 8};  // z: spin of a spin-batched launch
  }
  __device__ __forceinline__ void store(const Row& r, const Col& c, double c0, double c1) const {
    if (r.off < 0 || c.base == nullptr) return;
    *reinterpret_cast<double2*>(c.base + r.off) = make_double2(c0, c1);
  }
};

// ----------------------------------------------------------------------------
// TMA-fed version: persistent CTAs, one producer warp streaming both operands
// into a shared-memory ring with 1-D bulk copies (cp.async.bulk -> SASS UBLKCP,
// completion on mbarriers), CWM x CWN consumer warps issuing DMMA from
// conflict-free 256 B fragment reads.  The fragment-major HBM layout makes every
// (tile, k-range) a contiguous run, so no tensor map is needed.
// ----------------------------------------------------------------------------
constexpr int GT_STAGES = 3;  // ring depth
constexpr int GT_KS = 8;      // k-steps per stage (2 KB per tile row and stage)

template <int WM, int WN, int CWM, int CWN>
constexpr size_t gemm_tma_smem_bytes() {
  return (size_t)GT_STAGES * (WM * CWM + WN * CWN) * GT_KS * 32 * sizeof(double) + 2 * GT_STAGES * 8 + 128;
}

// Register rebalancing for the 8-consumer-warp configurations: the SM sub-partition that hosts the
// producer warp next to two consumer warps caps every thread at 168 registers (16384 / 3 warps),
// which spills the 4 x 8 accumulator blocks.  Launching a full producer warpgroup (12 warps) and
// moving registers with setmaxnreg (producer group 168 -> 40, consumers 168 -> 232; the increase
// 8 * 64 equals what the 4 producer-group warps release) removes the spills.
constexpr int GT_REGS_PRODUCER = 40, GT_REGS_CONSUMER = 232;
template <int NCW>
constexpr int gemm_tma_threads() { return NCW == 8 ? 12 * 32 : (NCW + 1) * 32; }

template <int WM, int WN, int CWM, int CWN, class Epi>
__global__ void __launch_bounds__(gemm_tma_threads<CWM * CWN>(), 1)
    gemm_tma_kernel(GemmArgs a, Epi epi, int tiles_m, int tiles_n, int batch) {
  constexpr int TM = WM * CWM, TN = WN * CWN, NCW = CWM * CWN;
  constexpr bool REB = NCW == 8;
  constexpr int A_STAGE = TM * GT_KS * 32, B_STAGE = TN * GT_KS * 32;  // doubles
  extern __shared__ __align__(128) double gt_smem[];
  double* As = gt_smem;
  double* Bs = gt_smem + GT_STAGES * A_STAGE;
  uint64_t* full = reinterpret_cast<uint64_t*>(Bs + GT_STAGES * B_STAGE);
  uint64_t* empty = full + GT_STAGES;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
#pragma unroll
    for (int s = 0; s < GT_STAGES; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], NCW * kReleaseArrivals);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int per_z = tiles_m * tiles_n;
  const int ntiles = per_z * batch;
  const int nkstage = (a.KS + GT_KS - 1) / GT_KS;

  if (warp >= NCW) {
    if constexpr (REB) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(GT_REGS_PRODUCER));
    if (warp != NCW) return;
    // ---------------- producer: lane L streams row L of the stage (A rows, then B rows) -----------
    static_assert(TM + TN <= 32, "one producer lane per tile row");
    unsigned it = 0;
    for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
      const int z = tile / per_z, r = tile % per_z;
      const int mt0 = (r % tiles_m) * TM, nt0 = (r / tiles_m) * TN;
      const int rows_m = min(TM, a.MTiles - mt0), rows_n = min(TN, a.NTiles - nt0);
      const double* src = nullptr;
      int dst_off = 0;
      bool isA = lane < TM;
      if (isA) {
        if (lane < rows_m) src = a.A + (size_t)z * a.strideAz + (size_t)(mt0 + lane) * a.KS * 32;
        dst_off = lane * GT_KS * 32;
      } else if (lane < TM + TN) {
        const int j = lane - TM;
        if (j < rows_n) {
          const int nt = nt0 + j;
          src = a.B + (size_t)z * a.strideBz + (size_t)(nt / a.ntInner) * a.strideBO +
                (size_t)(nt % a.ntInner) * a.strideBI;
        }
        dst_off = j * GT_KS * 32;
      }
      for (int ks = 0; ks < nkstage; ++ks, ++it) {
        const unsigned s = it % GT_STAGES, ph = (it / GT_STAGES) & 1u;
        mbar_wait(&empty[s], ph ^ 1u);
        const int k0 = ks * GT_KS, nk = min(GT_KS, a.KS - k0);
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
  if constexpr (REB) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(GT_REGS_CONSUMER));
  const int wm = warp % CWM, wn = warp / CWM;
  const int g = lane >> 2, t = lane & 3;
  const int boff = b_lane_offset(lane);
  unsigned it = 0;
  for (int tile = blockIdx.x; tile < ntiles; tile += gridDim.x) {
    const int z = tile / per_z, r = tile % per_z;
    const int mt0 = (r % tiles_m) * TM + wm * WM, nt0 = (r / tiles_m) * TN + wn * WN;
    double acc[WM][WN][2];
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
    for (int ks = 0; ks < nkstage; ++ks, ++it) {
      const unsigned s = it % GT_STAGES, ph = (it / GT_STAGES) & 1u;
      mbar_wait(&full[s], ph);
      const double* as = As + (size_t)s * A_STAGE + (size_t)wm * WM * GT_KS * 32 + lane;
      const double* bs = Bs + (size_t)s * B_STAGE + (size_t)wn * WN * GT_KS * 32 + boff;
      const int nk = min(GT_KS, a.KS - ks * GT_KS);
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
    typename Epi::Row er[WM];
    typename Epi::Col ec[WN];
#pragma unroll
    for (int i = 0; i < WM; ++i) er[i] = epi.row(min(mt0 + i, a.MTiles - 1), g);
#pragma unroll
    for (int j = 0; j < WN; ++j) ec[j] = epi.col(min(nt0 + j, a.NTiles - 1), z, t);
#pragma unroll
    for (int i = 0; i < WM; ++i) {
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        if (mt0 + i < a.MTiles && nt0 + j < a.NTiles) epi.store(er[i], ec[j], acc[i][j][0], acc[i][j][1]);
      }
    }
  }
}

template <int WM, int WN, int CWM, int CWN, class Epi>
inline cudaError_t launch_gemm_tma(const GemmArgs& a, const Epi& epi, int batch, int sm_count,
                                   cudaStream_t st) {
  constexpr int TM = WM * CWM, TN = WN * CWN;
  const int tiles_m = (a.MTiles + TM - 1) / TM, tiles_n = (a.NTiles + TN - 1) / TN;
  const int ntiles = tiles_m * tiles_n * batch;
  const size_t smem = gemm_tma_smem_bytes<WM, WN, CWM, CWN>();
  auto kern = gemm_tma_kernel<WM, WN, CWM, CWN, Epi>;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  const int grid = ntiles < sm_count ? ntiles : sm_count;
  kern<<<grid, gemm_tma_threads<CWM * CWN>(), smem, st>>>(a, epi, tiles_m, tiles_n, batch);
  return cudaGetLastError();
}

}  // namespace pxb

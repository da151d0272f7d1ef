// Taylor expansion exp(VHS) phi, persistent TMA-fed version (same polynomial and Horner
// evaluation as pxb_taylor.cuh; propagation/continuous.py:82-111,169-171).
//
// One CTA per SM, looping over (walker, orbital chunk) items:
//   * a producer warp streams the walker's VHS (A operand, written in fragment order by the VHS
//     GEMM epilogue) through a shared-memory ring with 1-D bulk copies (cp.async.bulk ->
//     UBLKCP, completion on mbarriers).  The stream is continuous over the Taylor orders and
//     over items, so the DMMA pipe never waits for a cold start at an order boundary;
//   * 8 consumer warps (2 per SM sub-partition) own rectangles of the (m-tile, n-tile) grid chosen
//     so that every sub-partition gets the same number of DMMAs: the 4 m-groups and 2 n-groups
//     are paired big-with-small (e.g. M=108, 42 orbitals: 14 x 11 tiles -> 4x6+3x5 = 39 and
//     3x6+4x5 = 38 tile products per sub-partition);
//   * the iterate S_n lives in shared memory in B-fragment order; with two buffers (when they
//     fit) an order needs a single consumer barrier.
#pragma once
#include "pxb_common.cuh"
#include "pxb_taylor.cuh"

namespace pxb {

#ifndef T2_KS_OVERRIDE
constexpr int T2_KS = 2;  // k-steps per ring stage
#else
constexpr int T2_KS = T2_KS_OVERRIDE;
#endif
// NG column groups of 4 consumer warps (one warp per SM sub-partition) + one producer warpgroup of
// which one warp works.  The CTA is launched with 65536 / threads registers per thread; setmaxnreg
// then moves registers from the producer warpgroup to the consumers.  setmaxnreg.inc can only take
// what setmaxnreg.dec released inside the same CTA (it blocks otherwise), so with R0 registers at
// launch  consumers * (Rc - R0) <= 4 * (R0 - Rp):
//   NG = 2: R0 = 168 (384 threads), Rp = 40, Rc = 232: a 7 x 6 tile block of accumulators fits
//   NG = 4: R0 =  96 (640 threads), Rp = 24, Rc = 112: 4 x 3 fits
constexpr int T2_MAXG = 4;
template <int NG>
struct T2Cfg {
  static constexpr int consumers = 4 * NG;
  static constexpr int threads = (consumers + 4) * 32;
  static constexpr int regs_producer = NG == 2 ? 40 : 24;
  static constexpr int regs_consumer = NG == 2 ? 232 : 112;
  static_assert(consumers * (regs_consumer - (65536 / threads) / 8 * 8) <= 4 * ((65536 / threads) / 8 * 8 - regs_producer),
                "setmaxnreg.inc would block: more registers requested than the producer warpgroup releases");
};

struct Taylor2Args {
  const double* VF;
  double* phi;
  const int* active;
  Dims d;
  int ochunk, nchunks;
  int NT;           // n-tiles per item (orbital slots / 4)
  int nbuf;         // 1 or 2 iterate buffers
  int nstage;       // ring depth
  int m_off[5];                 // m-group boundaries (4 groups)
  int n_off[T2_MAXG + 1];       // n-group (column group) boundaries
  int mperm[T2_MAXG][4];        // m-group of the warp of column group g on sub-partition s
  int dbg;                      // timing experiments only (PXB_TAYLOR_DBG): 1 no epilogue/barrier, 2 no phi reload
};

// The column groups (warps 4g .. 4g+3) own disjoint orbital columns, i.e. independent Taylor
// recursions: each synchronises on its own named barrier and they only meet at the VHS ring, so
// one group's epilogue / barrier wait overlaps the other groups' DMMA streams on every sub-partition.
__device__ __forceinline__ void bar_sync_group(int ng) {
  asm volatile("bar.sync %0, 128;" ::"r"(1 + ng) : "memory");
}

__device__ __forceinline__ double flip_sign_if(double v, unsigned mask) {
  return __hiloint2double(__double2hiint(v) ^ (int)mask, __double2loint(v));
}

inline size_t taylor2_smem_bytes(const Dims& d, int NT, int nbuf, int nstage) {
  return ((size_t)nbuf * d.KC * NT * 32 + (size_t)nstage * d.MT * T2_KS * 64) * sizeof(double) +
         2 * (size_t)nstage * 8 + 128;
}

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// S_order = phi columns o0..o0+no of walker (wg, wl), zero padded to 4*NT orbitals, into Tl
// (B-fragment order).  ASYNC: 16-byte cp.async copies (no registers, completes in the background).
template <bool ASYNC>
__device__ __forceinline__ void taylor2_load_tile(const Taylor2Args& a, double* Tl, int wg, int wl, int o0, int no,
                                                  int gtid, int nt0, int ntn) {
  const Dims& d = a.d;
  const int NT = a.NT;
  constexpr int GT = 128;                // threads of a warp group
  constexpr int U = 4;                   // independent loads in flight per thread (synchronous path)
  const int total = d.KC * ntn * 16;
  // the n-tiles [nt0, nt0 + ntn) of this warp group, by its 128 threads
  for (int base = gtid; base < total; base += GT * U) {
    double2 v[U];
    double2* dst[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int idx = base + u * GT;
      dst[u] = nullptr;
      v[u] = make_double2(0.0, 0.0);
      if (idx < total) {
        const int oo = idx & 3, tt = (idx >> 2) & 3, r = idx >> 4;
        const int nt = nt0 + r % ntn, kc = r / ntn;
        const int ol = 4 * nt + oo;
        dst[u] = reinterpret_cast<double2*>(Tl + ((size_t)kc * NT + nt) * 32 + tb_off(tt, 2 * oo));
        const double* src = a.phi + (((size_t)wg * d.ne + o0 + ol) * d.KC + kc) * 32 + wl * 8 + tt * 2;
        if (ol < no) {
          if (ASYNC) {
            cp_async_16(dst[u], src);
            dst[u] = nullptr;
          } else {
            v[u] = *reinterpret_cast<const double2*>(src);
          }
        }
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (dst[u] != nullptr) *dst[u] = v[u];
  }
}

// all Taylor orders of one item for a warp owning wm x wn tiles at (m0, n0).
// Horner step  S_{n-1} = phi + (VHS S_n) / n  with phi folded into the accumulators:
// acc starts at n * phi (loaded while the previous epilogue drains), so the epilogue is a
// multiply by 1/n and a shared-memory store with no global round trip inside it.
template <int WM, int WN>
__device__ __forceinline__ void taylor2_orders(const Taylor2Args& a, double* Tbuf, const double* ring,
                                               uint64_t* full, uint64_t* empty, unsigned& itc, int& cur,
                                               int m0, int n0, int wg, int wl, int o0, int no, int lane,
                                               int nwg, int nwl, int no0, int nno, int gtid, int ng) {
  const Dims& d = a.d;
  const int g = lane >> 2, t = lane & 3;
  const int NT = a.NT;
  const size_t tsz = (size_t)d.KC * NT * 32;
  const int boff = tb_off(t, g), boffp = tb_off(t, g ^ 1);
  const unsigned smask = (g & 1) ? 0u : 0x80000000u;  // (i B)^: (re, im) -> (-im, re)
  const int stage_doubles = d.MT * T2_KS * 64;
  const int nks = (d.KC + T2_KS - 1) / T2_KS;

  double acc[WM > 0 ? WM : 1][WN > 0 ? WN : 1][2];
  // acc = scale * phi (this lane's C-fragment elements)
  auto load_phi = [&](double scale) {
#pragma unroll
    for (int i = 0; i < WM; ++i) {
      const int p = 8 * (m0 + i) + g;
      const int kc2 = p >> 2, t2 = p & 3;
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        const int ol = 4 * (n0 + j) + t;
        double2 p0 = make_double2(0.0, 0.0);
        if (kc2 < d.KC && ol < no)
          p0 = *reinterpret_cast<const double2*>(a.phi + (((size_t)wg * d.ne + o0 + ol) * d.KC + kc2) * 32 +
                                                 wl * 8 + t2 * 2);
        acc[i][j][0] = scale * p0.x;
        acc[i][j][1] = scale * p0.y;
      }
    }
  };
  load_phi((double)d.exp_order);

  for (int n = d.exp_order; n >= 1; --n) {
    const double* Tcur = Tbuf + (size_t)cur * tsz;
    // during the last order the iterate buffer nobody reads receives the next item's phi tile
    if (n == 1 && a.nbuf == 2 && nwg >= 0)
      taylor2_load_tile<true>(a, Tbuf + (size_t)(cur ^ 1) * tsz, nwg, nwl, no0, nno, gtid, a.n_off[ng],
                              a.n_off[ng + 1] - a.n_off[ng]);
    for (int ks = 0; ks < nks; ++ks, ++itc) {
      const unsigned s = itc % (unsigned)a.nstage, ph = (itc / (unsigned)a.nstage) & 1u;
      mbar_wait(&full[s], ph);
      const int nk = min(T2_KS, d.KC - ks * T2_KS);
      if (WM > 0 && WN > 0) {
        const double* as = ring + (size_t)s * stage_doubles + (size_t)m0 * T2_KS * 64 + lane;
#pragma unroll
        for (int kk = 0; kk < T2_KS; ++kk) {
          if (kk < nk) {
            const double* Tk = Tcur + ((size_t)(ks * T2_KS + kk) * NT + n0) * 32;
            double ar[WM > 0 ? WM : 1], ai[WM > 0 ? WM : 1], b[WN > 0 ? WN : 1], bq[WN > 0 ? WN : 1];
#pragma unroll
            for (int i = 0; i < WM; ++i) {
              ar[i] = as[(i * T2_KS + kk) * 64];
              ai[i] = as[(i * T2_KS + kk) * 64 + 32];
            }
#pragma unroll
            for (int j = 0; j < WN; ++j) {
              b[j] = Tk[j * 32 + boff];
              bq[j] = flip_sign_if(Tk[j * 32 + boffp], smask);
            }
#pragma unroll
            for (int i = 0; i < WM; ++i)
#pragma unroll
              for (int j = 0; j < WN; ++j) dmma(acc[i][j][0], acc[i][j][1], ar[i], b[j]);
#pragma unroll
            for (int i = 0; i < WM; ++i)
#pragma unroll
              for (int j = 0; j < WN; ++j) dmma(acc[i][j][0], acc[i][j][1], ai[i], bq[j]);
          }
        }
      }
      ring_release(&empty[s], lane);
    }
    // S_{n-1} = (n phi + VHS S_n) / n  -> the other iterate buffer (or global for n == 1).
    // The reference divides by n (Temp = VHS.dot(Temp) / n); multiplying by the correctly rounded
    // reciprocal differs by at most one ulp per element.
    const double rn = 1.0 / (double)n;
    if (a.dbg & 1) continue;
    if (a.nbuf == 1) bar_sync_group(ng);  // the group has finished reading S_n
    double* Tnext = Tbuf + (size_t)(a.nbuf == 2 ? (cur ^ 1) : 0) * tsz;
#pragma unroll
    for (int i = 0; i < WM; ++i) {
      const int p = 8 * (m0 + i) + g;
      const int kc2 = p >> 2, t2 = p & 3;
      if (kc2 < d.KC) {
#pragma unroll
        for (int j = 0; j < WN; ++j) {
          const int ol = 4 * (n0 + j) + t;
          const double2 v = make_double2(acc[i][j][0] * rn, acc[i][j][1] * rn);
          if (n > 1) {
            *reinterpret_cast<double2*>(Tnext + ((size_t)kc2 * NT + n0 + j) * 32 + tb_off(t2, 2 * t)) = v;
          } else if (ol < no) {
            *reinterpret_cast<double2*>(a.phi + (((size_t)wg * d.ne + o0 + ol) * d.KC + kc2) * 32 + wl * 8 +
                                        t2 * 2) = v;
          }
        }
      }
    }
    if (n > 1) {
      if (!(a.dbg & 2)) load_phi((double)(n - 1));  // in flight across the barrier
      bar_sync_group(ng);         // S_{n-1} complete
      if (a.nbuf == 2) cur ^= 1;
    }
  }
}

// WMX = ceil(MT / 4), WNX = ceil(NT / 2): the largest warp rectangle; smaller groups use WMX-1 / WNX-1
template <int WMX, int WNX, int NG>
__global__ void __launch_bounds__(T2Cfg<NG>::threads, 1) taylor2_kernel(Taylor2Args a) {
  constexpr int T2_CONSUMERS = T2Cfg<NG>::consumers;
  extern __shared__ __align__(128) double t2_smem[];
  const Dims& d = a.d;
  const int NT = a.NT;
  const size_t tsz = (size_t)d.KC * NT * 32;
  double* Tbuf = t2_smem;
  double* ring = t2_smem + (size_t)a.nbuf * tsz;
  const int stage_doubles = d.MT * T2_KS * 64;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)a.nstage * stage_doubles);
  uint64_t* empty = full + a.nstage;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < a.nstage; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], T2_CONSUMERS * kReleaseArrivals);
    }
    fence_barrier_init();
  }
  __syncthreads();
  const int nitems = d.W * a.nchunks;
  const int nks = (d.KC + T2_KS - 1) / T2_KS;

  if (warp >= T2_CONSUMERS) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T2Cfg<NG>::regs_producer));
    if (warp != T2_CONSUMERS) return;
    // ---------------- producer: lane mt streams m-tile mt of the walker's VHS ----------------
    unsigned itc = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int w = item / a.nchunks;
      if (a.active != nullptr && a.active[w] == 0) continue;
      const double* src = a.VF + (size_t)w * vf_walker(d) + (size_t)lane * d.KC * 64;
      for (int n = 0; n < d.exp_order; ++n) {
        for (int ks = 0; ks < nks; ++ks, ++itc) {
          const unsigned s = itc % (unsigned)a.nstage, ph = (itc / (unsigned)a.nstage) & 1u;
          mbar_wait(&empty[s], ph ^ 1u);
          const int nk = min(T2_KS, d.KC - ks * T2_KS);
          const unsigned bytes = (unsigned)nk * 512u;
          if (lane == 0) mbar_expect_tx(&full[s], (unsigned)d.MT * bytes);
          __syncwarp();
          if (lane < d.MT)
            tma_bulk_g2s(ring + (size_t)s * stage_doubles + (size_t)lane * T2_KS * 64,
                         src + (size_t)ks * T2_KS * 64, bytes, &full[s]);
        }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(T2Cfg<NG>::regs_consumer));
  // warp w: column group w / 4 on sub-partition w % 4; the host permutes the m-groups per column
  // group so that every sub-partition gets the same number of tile products
  const int ng = warp >> 2, mg = a.mperm[ng][warp & 3];
  const int m0 = a.m_off[mg], wm = a.m_off[mg + 1] - m0;
  const int n0 = a.n_off[ng], wn = a.n_off[ng + 1] - n0;
  unsigned itc = 0;
  int cur = 0;
  const int gtid = tid & 127;  // thread index inside the warp group
  // first active item of this CTA
  auto next_active = [&](int item) {
    while (item < nitems && a.active != nullptr && a.active[item / a.nchunks] == 0) item += gridDim.x;
    return item;
  };
  int item = next_active(blockIdx.x);
  bool prefetched = false;
  while (item < nitems) {
    const int w = item / a.nchunks, chunk = item % a.nchunks;
    const int o0 = chunk * a.ochunk;
    const int no = min(a.ochunk, d.ne - o0);
    const int wg = w >> 2, wl = w & 3;
    const int nitem = next_active(item + gridDim.x);
    int nwg = -1, nwl = 0, no0 = 0, nno = 0;
    if (nitem < nitems) {
      const int w2 = nitem / a.nchunks, c2 = nitem % a.nchunks;
      nwg = w2 >> 2;
      nwl = w2 & 3;
      no0 = c2 * a.ochunk;
      nno = min(a.ochunk, d.ne - no0);
    }
    // S_order = phi tile of this item in the iterate buffer nobody reads
    const int ld = a.nbuf == 2 ? (cur ^ 1) : 0;
    if (prefetched) {
      cp_async_wait_all();  // issued during the previous item's last order
    } else {
      if (a.nbuf == 1) bar_sync_group(ng);
      taylor2_load_tile<false>(a, Tbuf + (size_t)ld * tsz, wg, wl, o0, no, gtid, n0, wn);
    }
    bar_sync_group(ng);
    cur = ld;
    prefetched = a.nbuf == 2 && nwg >= 0;
#define PXB_T2_CASE(WM_, WN_)                                                                              \
  taylor2_orders<WM_, WN_>(a, Tbuf, ring, full, empty, itc, cur, m0, n0, wg, wl, o0, no, lane, nwg, nwl, no0, \
                           nno, gtid, ng)
    if (wm == WMX && wn == WNX) PXB_T2_CASE(WMX, WNX);
    else if (wm == WMX && wn == WNX - 1) PXB_T2_CASE(WMX, (WNX - 1));
    else if (wm == WMX - 1 && wn == WNX) PXB_T2_CASE((WMX - 1), WNX);
    else if (wm == WMX - 1 && wn == WNX - 1) PXB_T2_CASE((WMX - 1), (WNX - 1));
    else __trap();
#undef PXB_T2_CASE
    item = nitem;
  }
}

}  // namespace pxb

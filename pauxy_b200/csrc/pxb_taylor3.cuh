// Taylor expansion exp(VHS) phi, third version (propagation/continuous.py:82-111,169-171): the
// persistent TMA-fed structure of pxb_taylor2.cuh with two changes.
//
// 1. Three real products per complex product.  With the iterate stored PLANAR (an n-tile is 8
//    orbitals of one plane, real | imaginary) the complex product C = A B becomes
//        P1 = Ar Br,   P2 = Ai Bi,   P3 = (Ar + Ai)(Br + Bi)
//        Re C = P1 - P2,   Im C = P3 - P1 - P2
//    i.e. 3 DMMAs per (8 rows x 8 orbitals x 4 k) instead of 4.  The sums Ar + Ai and Br + Bi are
//    formed in registers from the fragments (wm + wn DADDs per k-step against 3 wm wn DMMAs), so
//    neither the VHS stream nor the shared-memory footprint grows.  Same polynomial, same Horner
//    evaluation as before; the rounding differs at the 1e-16 level (normwise bound of the 3M
//    product), which the parity tests cover at their unchanged 1e-11 bar.
// 2. The item's phi tile stays in shared memory (planar) next to ONE iterate buffer that is updated
//    in place.  The accumulators of an order start at zero and the Horner "+ phi" is added in the
//    epilogue from shared memory, so no global load sits between a group barrier and the first
//    DMMA of the next order (re-reading phi from L2 into the accumulators cost 0.5 ms per step at
//    c4).  Each consumer warp owns a wm x wn block of (m-tile, n8-tile) pairs with all three
//    accumulator sets in registers (8 consumer warps at 232 registers after setmaxnreg).
//
// The ring / producer side, the per-column-group named barriers and the balanced assignment of
// warp rectangles to SM sub-partitions are those of taylor2_kernel.
#pragma once
#include "pxb_common.cuh"
#include "pxb_taylor2.cuh"

namespace pxb {

struct Taylor3Args {
  const double* VF;
  double* phi;
  const int* active;
  Dims d;
  int ochunk, nchunks;  // orbitals per item (multiple of 8), items per walker
  int NT8;              // n8-tiles per item
  int S;                // doubles per kc row of an iterate buffer: NT8 * 64 + 4 (skewed rows)
  int nstage;           // ring depth
  int nbuf;             // tile buffers: iterate + 1 phi tile (2) or iterate + 2 alternating phi tiles (3)
  int m_off[5];         // m-group boundaries (MG = 4 or 2 groups)
  int n_off[7];         // column-group boundaries (NG = 1, 2, 3 or 6 groups)
  int mperm[6][4];      // m-group of the k-th warp of column group g
  int dbg;              // timing experiments (PXB_EXPERIMENTS builds): 1 no epilogue / barriers, 2 no ring hand-shake, 4 no phi reload
};

// NG column groups of 4 consumer warps (one per SM sub-partition) + the producer warpgroup.
//   NG = 2:  8 consumer warps at 232 registers, warp rectangles up to 4 x 3 tile pairs
//   NG = 3: 12 consumer warps at 160 registers, up to 4 x 2: three warps per sub-partition cover
//           each other's epilogues better and nothing spills, at the price of an uneven split of
//           14 m-tiles x 3 groups over 4 sub-partitions (11 : 10)
//   NG = 1:  4 consumer warps, TWO CTAs per SM (two walkers in flight, each with its own VHS ring): shapes
//           with at most 16 orbital columns (c3: 14), where splitting the columns would leave a warp
//           6 DMMAs per k-step; 128 registers, no rebalancing
//   NG = 6, MG = 2: 12 consumer warps in six column groups of TWO warps (7 x 1 tile pairs each at c4:
//           14 m-tiles x 6 column tiles split evenly, 21 tile pairs on every sub-partition instead of
//           22 : 20, and a group rendez-vous between two warps instead of four)
template <int NG, int MG = 4>
struct T3Cfg {
  static constexpr int consumers = MG * NG;
  static constexpr int threads = (consumers + 4) * 32;
  static constexpr int ctas_per_sm = NG == 1 ? 2 : 1;
  static constexpr bool rebalance = NG > 1;  // setmaxnreg: producer warpgroup -> consumers
  static constexpr int regs_producer = consumers == 8 ? 40 : 24;
  static constexpr int regs_consumer = consumers == 8 ? 232 : 160;
  static_assert(!rebalance ||
                    consumers * (regs_consumer - (65536 / threads) / 8 * 8) <= 4 * ((65536 / threads) / 8 * 8 - regs_producer),
                "setmaxnreg.inc would block: more registers requested than the producer warpgroup releases");
};
constexpr int T3_MIN_STAGES2 = 3;  // fewest ring stages accepted with two iterate buffers

inline size_t taylor3_smem_bytes(const Dims& d, int NT8, int nbuf, int nstage) {
  const size_t S = (size_t)NT8 * 64 + 4;
  return ((size_t)nbuf * d.KC * S + (size_t)nstage * d.MT * T2_KS * 64) * sizeof(double) + 2 * (size_t)nstage * 8 + 12 * 8 + 128;
}

__device__ __forceinline__ void bar_sync_threads(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void cp_async_8(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_u32(smem_dst)), "l"(gmem_src) : "memory");
}

// phi columns o0 .. o0 + no of walker (wg, wl) -> planar B fragments of the n8-tiles [nt0, nt0 + ntn)
// of buffer T (element (k = p & 3, n = ol & 7) of fragment (kc = p >> 2, nt = ol >> 3, plane) sits at
// 4 n + k), by the 128 (or 64) threads of a column group; 8-byte cp.async copies de-interleave (re, im) on
// the fly and complete in the background, padding orbitals are zero-filled with plain stores
__device__ __forceinline__ void taylor3_load_tile(const Taylor3Args& a, double* T, int wg, int wl, int o0, int no,
                                                  int gtid, int nt0, int ntn, int gthreads = 128) {
  const Dims& d = a.d;
  const int tc = gtid & 7, tt = tc >> 1, c = tc & 1;
  const int ncol = 8 * ntn;
  for (int kc = 0; kc < d.KC; ++kc) {
    double* row = T + (size_t)kc * a.S + c * 32 + tt;
    for (int oll = gtid >> 3; oll < ncol; oll += gthreads >> 3) {
      const int ol = 8 * nt0 + oll;
      double* dst = row + (ol >> 3) * 64 + 4 * (ol & 7);
      if (ol < no)
        cp_async_8(dst, a.phi + (((size_t)wg * d.ne + o0 + ol) * d.KC + kc) * 32 + wl * 8 + tc);
      else
        *dst = 0.0;
    }
  }
}

struct T3Frag {  // fragments of one k-step for a WM x WN warp block (WM <= 7, WN <= 3)
  double ar[7], ai[7], br[3], bi[3];
};
static_assert(T2_KS == 2, "the k loop of taylor3_orders is written for two k-steps per ring stage");

// shared-memory accesses through 32-bit shared-space addresses (one register per base address,
// immediate offsets): the consumer loop lives at the edge of the 232-register budget
__device__ __forceinline__ double lds_f64(uint32_t addr) {
  double v;
  asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts_f64(uint32_t addr, double v) {
  asm volatile("st.shared.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void mbar_wait_u32(uint32_t bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void ring_release_u32(uint32_t bar, int lane) {
#ifdef PXB_RACECHECK_STRICT
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
#else
  __syncwarp();
  if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
#endif
}

template <int WM, int WN>
__device__ __forceinline__ void t3_load(T3Frag& f, uint32_t ap, uint32_t bp) {
#pragma unroll
  for (int i = 0; i < WM; ++i) {
    f.ar[i] = lds_f64(ap + i * (T2_KS * 64 * 8));
    f.ai[i] = lds_f64(ap + i * (T2_KS * 64 * 8) + 256);
  }
#pragma unroll
  for (int j = 0; j < WN; ++j) {
    f.br[j] = lds_f64(bp + j * 512);
    f.bi[j] = lds_f64(bp + j * 512 + 256);
  }
}

// one k-step: P1 += Ar Br, P2 += Ai Bi, P3 += (Ar + Ai)(Br + Bi); 3 WM WN independent DMMAs
template <int WM, int WN>
__device__ __forceinline__ void t3_kstep(double (&P1)[WM][WN][2], double (&P2)[WM][WN][2], double (&P3)[WM][WN][2],
                                         uint32_t ap, uint32_t bp) {
  T3Frag f;
  t3_load<WM, WN>(f, ap, bp);
  double as[WM], bs[WN];
#pragma unroll
  for (int i = 0; i < WM; ++i) as[i] = f.ar[i] + f.ai[i];
#pragma unroll
  for (int j = 0; j < WN; ++j) bs[j] = f.br[j] + f.bi[j];
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) dmma(P1[i][j][0], P1[i][j][1], f.ar[i], f.br[j]);
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) dmma(P2[i][j][0], P2[i][j][1], f.ai[i], f.bi[j]);
#pragma unroll
  for (int i = 0; i < WM; ++i)
#pragma unroll
    for (int j = 0; j < WN; ++j) dmma(P3[i][j][0], P3[i][j][1], as[i], bs[j]);
}

// All Taylor orders of one item for a warp owning WM x WN (m-tile, n8-tile) pairs at (m0, n0).
// Horner step S_{n-1} = phi + (VHS S_n) / n.  Shared memory holds the item's phi tile (planar,
// read-only for the whole item: it is S_N, and the "+ phi" of every epilogue) and ONE iterate
// buffer updated in place: the accumulators start at zero, so an order begins with DMMAs the moment
// its operand is complete (no loads in front of it), and the epilogue adds phi from shared memory.
template <int WM, int WN>
__device__ __forceinline__ void taylor3_orders(const Taylor3Args& a, uint32_t phib, uint32_t itb, uint32_t ring,
                                               uint32_t full, uint32_t empty, unsigned& rs, unsigned& rph,
                                               int m0, int n0, double* gphi, int no, int lane, int ng, uint32_t gbar,
                                               unsigned& gph) {
  const Dims& d = a.d;
  const int g = lane >> 2, t = lane & 3;
  const int Sb = a.S * 8;                        // bytes per kc row of a tile buffer
  const int stage_bytes = d.MT * T2_KS * 64 * 8;
  const int nks = (d.KC + T2_KS - 1) / T2_KS;
  const uint32_t a_off = (uint32_t)(m0 * (T2_KS * 64) + lane) * 8;  // this warp's A fragments inside a ring stage
  const uint32_t b_off = (uint32_t)(n0 * 64 + lane) * 8;            // its B fragments inside a kc row
  // C fragment element (row 8 mt + g, columns 8 nt + 2t + e) as B fragment element of a tile
  // buffer: kc = 2 mt + (g >> 2), position 4 (2t + e) + (g & 3)
  const uint32_t st_off = (uint32_t)((g >> 2) * a.S + 8 * t + (g & 3)) * 8 + (uint32_t)(2 * m0) * Sb + n0 * 512;

  const uint32_t pad_last = 2 * (m0 + WM - 1) + (g >> 2) >= d.KC ? (uint32_t)Sb : 0u;
  double P1[WM][WN][2], P2[WM][WN][2], P3[WM][WN][2];
  for (int n = d.exp_order; n >= 1; --n) {
#pragma unroll
    for (int i = 0; i < WM; ++i)
#pragma unroll
      for (int j = 0; j < WN; ++j)
#pragma unroll
        for (int e = 0; e < 2; ++e) P1[i][j][e] = P2[i][j][e] = P3[i][j][e] = 0.0;
    // ---- k loop: one ring stage (two k-steps) per iteration, "load fragments, 3 wm wn
    //      independent DMMAs" per k-step (tools/micro/dmma_loop_micro.cu: 93 % of the pipe with
    //      two warps per sub-partition when nothing spills) ----
    uint32_t bp = (n == d.exp_order ? phib : itb) + b_off;
#pragma unroll 1
    for (int ks = 0; ks < nks; ++ks) {
      if (!(a.dbg & 2)) mbar_wait_u32(full + rs * 8, rph);
      const uint32_t ap = ring + rs * stage_bytes + a_off;
      t3_kstep<WM, WN>(P1, P2, P3, ap, bp);
      bp += Sb;
      if (ks * T2_KS + 1 < d.KC) {
        t3_kstep<WM, WN>(P1, P2, P3, ap + 512, bp);
        bp += Sb;
      }
      if (!(a.dbg & 2)) ring_release_u32(empty + rs * 8, lane);
      if (++rs == (unsigned)a.nstage) {
        rs = 0;
        rph ^= 1u;
      }
    }
    if (a.dbg & 1) continue;
    // ---- S_{n-1} = phi + (VHS S_n) / n -> the iterate buffer in place (global for n == 1).  The
    // reference divides by n; multiplying by the correctly rounded reciprocal differs by at most
    // one ulp per element.
    const double rn = 1.0 / (double)n;
    // The group's two rendez-vous per order are split into "arrive" and "wait" (mbarriers, one
    // arrival per warp): a warp announces that it has finished READING S_n, does the epilogue
    // arithmetic (phi comes from the read-only phi tile), and only then waits for the others
    // before it overwrites its part of the iterate in place.
    const bool inplace = n > 1 && n < d.exp_order;
    if (inplace) {
      ring_release_u32(gbar, lane);  // __syncwarp + lane-0 arrive (every lane in the checker build)
    }
#pragma unroll
    for (int i = 0; i < WM; ++i) {
#pragma unroll
      for (int j = 0; j < WN; ++j) {
        // a padding row (kc row KC of the last m-tile when KC is odd; only the warp's last tile can
        // hold it) reads row KC - 1 of the phi tile instead of whatever lies behind the tile (the VHS
        // ring); its values are never stored
        const uint32_t so = st_off + (uint32_t)(2 * i) * Sb + j * 512 - (i == WM - 1 ? pad_last : 0u);
#pragma unroll
        for (int e = 0; e < 2; ++e) {
          const double re = (P1[i][j][e] - P2[i][j][e]) * rn + lds_f64(phib + so + e * 32);
          const double im = ((P3[i][j][e] - P1[i][j][e]) - P2[i][j][e]) * rn + lds_f64(phib + so + e * 32 + 256);
          P1[i][j][e] = re;
          P3[i][j][e] = im;
        }
      }
    }
    if (inplace) {
      mbar_wait_u32(gbar, gph & 1u);
      gph ^= 1u;
    }
#pragma unroll
    for (int i = 0; i < WM; ++i) {
      const int kc2 = 2 * (m0 + i) + (g >> 2);
      if (kc2 < d.KC) {
#pragma unroll
        for (int j = 0; j < WN; ++j) {
          const uint32_t so = st_off + (uint32_t)(2 * i) * Sb + j * 512;
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            if (n > 1) {
              sts_f64(itb + so + e * 32, P1[i][j][e]);
              sts_f64(itb + so + e * 32 + 256, P3[i][j][e]);
            } else {
              const int ol = 8 * (n0 + j) + 2 * t + e;
              if (ol < no)
                *reinterpret_cast<double2*>(gphi + ((size_t)ol * d.KC + kc2) * 32) = make_double2(P1[i][j][e], P3[i][j][e]);
            }
          }
        }
      }
    }
    if (n > 1) {  // S_{n-1} complete
      ring_release_u32(gbar + 8, lane);
      mbar_wait_u32(gbar + 8, (gph >> 1) & 1u);
      gph ^= 2u;
    }
  }
}

// WMX = ceil(MT / 4), WNX = ceil(NT8 / NG): the largest warp rectangle; smaller groups use WMX-1 / WNX-1
template <int WMX, int WNX, int NG, int MG = 4>
__global__ void __launch_bounds__(T3Cfg<NG, MG>::threads, T3Cfg<NG, MG>::ctas_per_sm) taylor3_kernel(Taylor3Args a) {
  constexpr int T3_CONSUMERS = T3Cfg<NG, MG>::consumers;
  extern __shared__ __align__(128) double t3_smem[];
  const Dims& d = a.d;
  const size_t tsz = (size_t)d.KC * a.S;
  double* Tbuf = t3_smem;            // [0]: iterate (in place), [1] and [2]: phi tiles (current / next item)
  double* ring = t3_smem + (size_t)a.nbuf * tsz;
  const int stage_doubles = d.MT * T2_KS * 64;
  uint64_t* full = reinterpret_cast<uint64_t*>(ring + (size_t)a.nstage * stage_doubles);
  uint64_t* empty = full + a.nstage;
  uint64_t* group_bar = empty + a.nstage;  // per column group: "finished reading", "iterate complete"
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) {
    for (int s = 0; s < a.nstage; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], T3_CONSUMERS * kReleaseArrivals);
    }
    for (int g2 = 0; g2 < 2 * NG; ++g2) mbar_init(&group_bar[g2], MG * kReleaseArrivals);
    fence_barrier_init();
  }
  __syncthreads();
  const int nitems = d.W * a.nchunks;
  const int nks = (d.KC + T2_KS - 1) / T2_KS;

  if (warp >= T3_CONSUMERS) {
    if constexpr (T3Cfg<NG, MG>::rebalance) asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(T3Cfg<NG, MG>::regs_producer));
    if (warp != T3_CONSUMERS || (a.dbg & 2)) return;
    // ---------------- producer: lane mt streams m-tile mt of the walker's VHS ----------------
    unsigned s = 0, ph = 0;
    for (int item = blockIdx.x; item < nitems; item += gridDim.x) {
      const int w = item / a.nchunks;
      if (a.active != nullptr && a.active[w] == 0) continue;
      const double* src = a.VF + (size_t)w * vf_walker(d) + (size_t)lane * d.KC * 64;
      for (int n = 0; n < d.exp_order; ++n) {
        for (int ks = 0; ks < nks; ++ks) {
          mbar_wait(&empty[s], ph ^ 1u);
          const int nk = min(T2_KS, d.KC - ks * T2_KS);
          const unsigned bytes = (unsigned)nk * 512u;
          if (lane == 0) mbar_expect_tx(&full[s], (unsigned)d.MT * bytes);
          __syncwarp();
          if (lane < d.MT)
            tma_bulk_g2s(ring + (size_t)s * stage_doubles + (size_t)lane * T2_KS * 64,
                         src + (size_t)ks * T2_KS * 64, bytes, &full[s]);
          if (++s == (unsigned)a.nstage) {
            s = 0;
            ph ^= 1u;
          }
        }
      }
    }
    return;
  }

  // ---------------- consumers ----------------
  if constexpr (T3Cfg<NG, MG>::rebalance) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(T3Cfg<NG, MG>::regs_consumer));
  const int ng = warp / MG, mg = a.mperm[ng][warp % MG];
  const int m0 = a.m_off[mg], wm = a.m_off[mg + 1] - m0;
  const int n0 = a.n_off[ng], wn = a.n_off[ng + 1] - n0;
  unsigned rs = 0, rph = 0;
  unsigned gph = 0;  // phases of this group's two barriers
  const uint32_t gbar = smem_u32(group_bar + 2 * ng);
  const int gtid = tid % (32 * MG);  // thread index inside the column group
  auto next_active = [&](int item) {
    while (item < nitems && a.active != nullptr && a.active[item / a.nchunks] == 0) item += gridDim.x;
    return item;
  };
  // phi tiles: with nbuf == 3 two of them alternate, the next item's is fetched (cp.async) while the
  // current item runs; with nbuf == 2 there is one and it is loaded at the start of every item
  int pcur = 1;
  int item = next_active(blockIdx.x);
  bool prefetched = false;
  while (item < nitems) {
    const int w = item / a.nchunks, chunk = item % a.nchunks;
    const int o0 = chunk * a.ochunk;
    const int no = min(a.ochunk, d.ne - o0);
    const int wg = w >> 2, wl = w & 3;
    const int nitem = next_active(item + gridDim.x);
    if (!prefetched) {
      if (a.nbuf == 2) bar_sync_threads(1 + ng, 32 * MG);  // the group has finished with the previous item's phi tile
      taylor3_load_tile(a, Tbuf + (size_t)pcur * tsz, wg, wl, o0, no, gtid, n0, wn, 32 * MG);
    }
    cp_async_wait_all();
    bar_sync_threads(1 + ng, 32 * MG);
    prefetched = false;
    if (a.nbuf == 3 && nitem < nitems) {  // next item's phi tile into the other phi buffer
      const int w2 = nitem / a.nchunks, c2 = nitem % a.nchunks;
      const int no0 = c2 * a.ochunk;
      taylor3_load_tile(a, Tbuf + (size_t)(pcur ^ 3) * tsz, w2 >> 2, w2 & 3, no0, min(a.ochunk, d.ne - no0), gtid, n0,
                        wn, 32 * MG);
      prefetched = true;
    }
    double* gphi = a.phi + ((size_t)wg * d.ne + o0) * d.KC * 32 + wl * 8 + ((lane >> 2) & 3) * 2;
    const uint32_t phib = smem_u32(Tbuf + (size_t)pcur * tsz), itb = smem_u32(Tbuf);
#define PXB_T3_CASE(WM_, WN_)                                                                                  \
  taylor3_orders<WM_, WN_>(a, phib, itb, smem_u32(ring), smem_u32(full), smem_u32(empty), rs, rph, m0, n0, gphi, no, \
                           lane, ng, gbar, gph)
    if (wm == WMX && wn == WNX) PXB_T3_CASE(WMX, WNX);
    else if (WNX > 1 && wm == WMX && wn == WNX - 1) PXB_T3_CASE(WMX, (WNX > 1 ? WNX - 1 : 1));
    else if (WMX > 1 && wm == WMX - 1 && wn == WNX) PXB_T3_CASE((WMX > 1 ? WMX - 1 : 1), WNX);
    else if (WMX > 1 && WNX > 1 && wm == WMX - 1 && wn == WNX - 1)
      PXB_T3_CASE((WMX > 1 ? WMX - 1 : 1), (WNX > 1 ? WNX - 1 : 1));
    else __trap();
#undef PXB_T3_CASE
    if (a.nbuf == 3) pcur ^= 3;
    item = nitem;
  }
}

}  // namespace pxb

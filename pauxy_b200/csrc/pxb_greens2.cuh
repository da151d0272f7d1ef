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
#include <type_traits>
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
  const int nc = (nmax + 7) / 8 * 8;
  // inverse [nmax][nld], colk [nmax] + 3 elements of slack (the A fragments of the last k-step read up
  // to 3 entries past a row), the pivot-row exchange buffer [2][nc] of the register path, piv [nc]
  // and its pivots with their reciprocals [2][nc]
  const int regpath = nmax <= 32 ? 4 * nc : 0;  // exchange buffer + pivots: only where gj_invert_regs runs
  return ((size_t)nmax * nld * sizeof(cplx) + (size_t)(nmax + 3) * sizeof(cplx) + (size_t)regpath * sizeof(cplx) +
          (size_t)nc * sizeof(int) + 15) / 16 * 16;
}

// Register-resident Gauss-Jordan inversion of the ns x ns matrix in A (ns <= NC <= 32), result back
// in A, with LAPACK's pivot choice (zgetrf: first row of largest |re| + |im| in the column).
//   - lane i owns row i of the matrix, zero-padded to NC columns: 2 NC doubles in registers; rows are never moved, every lane tracks the position its row has in the pivoted
//     order instead (pos), so ties are resolved exactly as with physical interchanges;
//   - each step works on column slot 0 and writes the updated row shifted down by one slot
//     (new[j-1] = old[j] - f p[j], new[NC-1] = the column of the inverse): the active column has a
//     static register index although the step loop is not unrolled; after the ns steps the column
//     of the inverse that pivot step j produced sits in slot j + NC - ns;
//   - pivot rows stay unscaled (their column-of-the-inverse entry is 1); a row is linear in its own
//     scale, so each lane multiplies its row by the reciprocal of its pivot once at the end;
//   - pivot search: two 32-bit REDUX.MAX over the bit pattern of the (non-negative) magnitudes and a
//     REDUX.MIN over the positions of the tied rows; the pivot row travels through a
//     double-buffered row in shared memory (one predicated STS + one broadcast LDS per element).
// sign, logdet: each lane takes log / phase of the pivot it supplied; warp tree reduction.
template <int NC>
__device__ __forceinline__ void gj_invert_regs(cplx* A, int lda, int ns, cplx* prow, int* piv, int lane, cplx& sign,
                                               double& logdet) {
  constexpr unsigned FULL = 0xffffffffu;
  double2 r[NC];
#pragma unroll
  for (int j = 0; j < NC; ++j) {
    r[j] = make_double2(0.0, 0.0);
    if (lane < ns && j < ns) r[j] = *reinterpret_cast<const double2*>(A + (size_t)lane * lda + j);
  }
  unsigned pos = lane < ns ? (unsigned)lane : 0xffu;
  bool done = lane >= ns;
  double2* pvk = reinterpret_cast<double2*>(prow) + 2 * NC;  // pivot of step k and its reciprocal: [k], [NC + k]
  bool odd = false;
  __syncwarp();
#pragma unroll 1
  for (int k = 0; k < ns; ++k) {
    const double v = fabs(r[0].x) + fabs(r[0].y);
    const unsigned vh = (unsigned)__double2hiint(v), vl = (unsigned)__double2loint(v);
    const unsigned kh = done ? 0u : vh + 1u;
    const unsigned mh = __reduce_max_sync(FULL, kh);
    const bool top = !done && kh == mh;
    const unsigned ml = __reduce_max_sync(FULL, top ? vl : 0u);
    const bool tie = top && vl == ml;
    const unsigned mp = __reduce_min_sync(FULL, tie ? pos : 0xffu);
    const bool is_piv = tie && pos == mp;
    if (mp != (unsigned)k) {  // interchange of positions k and mp (warp-uniform)
      odd = !odd;
      if (pos == (unsigned)k) pos = mp;
    }
    double2* pb = reinterpret_cast<double2*>(prow) + (k & 1) * NC;
    if (is_piv) {
      pos = (unsigned)k;
      if (true) piv[k] = lane;
#pragma unroll
      for (int j = 0; j < NC; ++j) pb[j] = r[j];
    }
    __syncwarp();
    const double2 pv = pb[0];
    // 1 / pv with one division: pv is scaled by the power of two that brings its larger component
    // into [1, 2) (built from the exponent bits, exact), so that a^2 + b^2 lies in [1, 8)
    const int eh = max(__double2hiint(pv.x) & 0x7ff00000, __double2hiint(pv.y) & 0x7ff00000);
    const double sc = __hiloint2double(0x7fe00000 - eh, 0);
    const double pa = pv.x * sc, pbi = pv.y * sc;
    const double inv = sc / (pa * pa + pbi * pbi);
    const double2 rp = make_double2(pa * inv, -pbi * inv);
    double2 f = make_double2(r[0].x * rp.x - r[0].y * rp.y, r[0].x * rp.y + r[0].y * rp.x);
    if (is_piv) {
      f = make_double2(0.0, 0.0);
      pvk[k] = pv;
      pvk[NC + k] = rp;
      done = true;
    }
#pragma unroll
    for (int j = 1; j < NC; ++j) {
      const double2 p = pb[j];
      r[j - 1].x = r[j].x - (f.x * p.x - f.y * p.y);
      r[j - 1].y = r[j].y - (f.x * p.y + f.y * p.x);
    }
    r[NC - 1] = is_piv ? make_double2(1.0, 0.0) : make_double2(-f.x, -f.y);
  }
  __syncwarp();  // piv[] complete; every lane is past its reads of A
  // the row of lane i was the pivot row of step pos
  const double2 mypv = lane < ns ? pvk[pos] : make_double2(1.0, 0.0);
  const double2 myrp = lane < ns ? pvk[NC + pos] : make_double2(1.0, 0.0);
  // row `pos` of the inverse; after ns steps the column of pivot step j sits in slot j + NC - ns
  if (lane < ns) {
    cplx* arow = A + (size_t)pos * lda;
    const int* pj = piv - (NC - ns);
#pragma unroll
    for (int j = 0; j < NC; ++j) {
      if (j >= NC - ns) {
        const double2 x = r[j];
        arow[pj[j]] = {x.x * myrp.x - x.y * myrp.y, x.x * myrp.y + x.y * myrp.x};
      }
    }
  }
  // det = (-1)^interchanges prod pivots
  const double au = hypot(mypv.x, mypv.y);
  double ld = log(au);
  cplx sg = {mypv.x / au, mypv.y / au};
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    ld += __shfl_xor_sync(FULL, ld, m);
    const cplx o = {__shfl_xor_sync(FULL, sg.re, m), __shfl_xor_sync(FULL, sg.im, m)};
    sg = cmul(sg, o);
  }
  sign = odd ? cplx{-sg.re, -sg.im} : sg;
  logdet = ld;
  __syncwarp();
}

// out[a][p] = sum_i A[a][i] in[p][i] for one (walker, spin) by one warp on DMMA: A is an ns x ns complex
// matrix, row-major with leading dimension lda in shared memory (readable up to 3 elements past each
// row), `in` / `out` point at the walker's slot of the first orbital row of the spin in the OF layout
// (base + wl * 8).  in == out is allowed: a chunk is read completely before it is stored.
// E1B: also er + i ei += sum_{a,p} h1rot[a][p] out[a][p] (this lane's share; h1rot at the spin's first row).
// NC: read `in` through the non-coherent path (only when nothing in this kernel writes it).
// LOWER: A is lower triangular (zero tiles are skipped; needs every k-step slot in use, KS == 2 NMT).
template <int NMT, bool E1B, bool NC, bool LOWER = false>
__device__ __forceinline__ void warp_apply_left(const cplx* A, int lda, int ns, const Dims& d, const double* in,
                                                double* out, const double2* h1rot, int lane, double& er, double& ei) {
  const int g = lane >> 2, t = lane & 3;
  // one basis chunk (n-tile of 4 basis functions) per
  //    iteration.  Everything that does not depend on the chunk is hoisted: A fragments are read
  //    from the row-major inverse with one base address per m-tile plus immediates (rows >= ns of the
  //    last m-tile are clamped to row ns - 1: a row of A only reaches the same row of C, and those
  //    rows are never stored; columns >= ns read the start of the next row or of colk, finite
  //    numbers, and meet B entries that are forced to zero); B / Theta / h1rot are addressed with
  //    pointers that advance by one chunk; only the LAST k-step can hold padding (4 (KS-1) + t >= ns).
  //    The B fragments of the next chunk and the h1rot entries of this one are requested before the
  //    DMMAs of the current chunk (16 warps per SM: latency has to be hidden inside the warp).
  const int KS = (ns + 3) >> 2;
  constexpr int KSM = 2 * NMT;  // >= ceil(ns / 4)
  const unsigned rowB = (unsigned)d.KC * 32u;
  const unsigned smask = (g & 1) ? 0u : 0x80000000u;  // (i B)^: (re, im) -> (-im, re)
  const bool last_pad = 4 * (KS - 1) + t >= ns;       // this lane's entry of the last k-step is padding
  const double* bp = in + (size_t)t * rowB + g;  // k-step 0, chunk 0
  const double* bl = in + (size_t)min(4 * (KS - 1) + t, ns - 1) * rowB + g;
  const unsigned brow4 = 4u * rowB;
  const double2* Am[NMT];  // rows >= ns are clamped to the last row (their C rows are not stored)
#pragma unroll
  for (int m = 0; m < NMT; ++m) Am[m] = reinterpret_cast<const double2*>(A + (size_t)min(8 * m + g, ns - 1) * lda + t);
  const int klast = 4 * (KS - 1);
  // up to 32 orbitals the A fragments stay in registers for all chunks (the shared-memory data path
  // is what bounds this kernel: 6 NMT^2 LDS.128 per chunk otherwise)
  constexpr bool AREG = NMT <= 4;
  double2 af[AREG ? NMT : 1][AREG ? KSM : 1];
  if constexpr (AREG) {
#pragma unroll
    for (int ks = 0; ks < KSM; ++ks)
#pragma unroll
      for (int m = 0; m < NMT; ++m)
        af[m][ks] = ks == KSM - 1 ? Am[m][klast] : ks < KS - 1 ? Am[m][4 * ks] : make_double2(0.0, 0.0);
  }
  double* tp = out + (size_t)g * rowB + t * 2;
  const double2* hp = E1B ? h1rot + (size_t)g * d.Mp + t : nullptr;
  const unsigned trow8 = 8u * rowB, hrow8 = 8u * (unsigned)d.Mp;
  bool rowok[NMT];
#pragma unroll
  for (int m = 0; m < NMT; ++m) rowok[m] = 8 * m + g < ns;
  // CH basis chunks per iteration: one when the A fragments sit in registers, two when they are read
  // from shared memory (each A fragment then serves both chunks)
  constexpr int CH = AREG ? 1 : 2;
  double bn[KSM][CH];
  // KFULL: every k-step slot is in use (KS == KSM, e.g. 21 orbitals in 6 k-steps): no guards
  auto chunks = [&](auto kfull_tag) {
  constexpr bool KFULL = decltype(kfull_tag)::value;
  auto load_b = [&](int pc0) {  // slot KSM - 1 holds k-step KS - 1, whatever KS is
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      // a chunk past the end (odd KC with CH = 2) repeats the last one; its results are not stored
      const int o = (q > 0 && pc0 + q >= d.KC) ? 0 : 32 * q;
#pragma unroll
      for (int ks = 0; ks < KSM - 1; ++ks)
        if (KFULL || ks < KS - 1) bn[ks][q] = NC ? ldg_nc(bp + ks * brow4 + o) : bp[ks * brow4 + o];
      bn[KSM - 1][q] = last_pad ? 0.0 : NC ? ldg_nc(bl + o) : bl[o];
    }
    bp += 32 * CH;
    bl += 32 * CH;
  };
  load_b(0);
  const int pz0 = d.M - t;  // entries with 4 pc >= pz0 are basis padding
  for (int pc = 0; pc < d.KC; pc += CH) {
    double b[KSM][CH];
#pragma unroll
    for (int ks = 0; ks < KSM; ++ks)
#pragma unroll
      for (int q = 0; q < CH; ++q) b[ks][q] = bn[ks][q];
    if (pc + CH < d.KC) load_b(pc + CH);
    double2 hq[NMT][CH];
#pragma unroll
    for (int q = 0; q < CH; ++q)
#pragma unroll
      for (int m = 0; m < NMT; ++m)
        hq[m][q] = E1B && rowok[m] && pc + q < d.KC ? hp[m * hrow8 + 4 * q] : make_double2(0.0, 0.0);
    if (E1B) hp += 4 * CH;
    double acc[NMT][CH][2];
#pragma unroll
    for (int m = 0; m < NMT; ++m)
#pragma unroll
      for (int q = 0; q < CH; ++q) acc[m][q][0] = acc[m][q][1] = 0.0;
#pragma unroll
    for (int ks = 0; ks < KSM; ++ks) {
      const bool lastk = ks == KSM - 1;
      if (KFULL || lastk || ks < KS - 1) {
        double bq[CH];
#pragma unroll
        for (int q = 0; q < CH; ++q) {
          const double o = __shfl_xor_sync(0xffffffffu, b[ks][q], 4);  // the other component of the same element
          bq[q] = __hiloint2double(__double2hiint(o) ^ (int)smask, __double2loint(o));
        }
#pragma unroll
        for (int m = 0; m < NMT; ++m) {  // m-tiles beyond ceil(ns / 8) (uneven spins) compute on clamped rows
          if (LOWER && KFULL && 4 * ks > 8 * m + 7) continue;  // A[8m..8m+7][4ks..] = 0
          double2 av;
          if constexpr (AREG) av = af[m][ks];
          else av = lastk ? Am[m][klast] : Am[m][4 * ks];
#pragma unroll
          for (int q = 0; q < CH; ++q) {
            dmma(acc[m][q][0], acc[m][q][1], av.x, b[ks][q]);
            dmma(acc[m][q][0], acc[m][q][1], av.y, bq[q]);
          }
        }
      }
    }
#pragma unroll
    for (int q = 0; q < CH; ++q) {
      if (pc + q >= d.KC) continue;
      const bool pz = 4 * (pc + q) >= pz0;
#pragma unroll
      for (int m = 0; m < NMT; ++m) {
        if (rowok[m]) {
          const double vr = pz ? 0.0 : acc[m][q][0], vi = pz ? 0.0 : acc[m][q][1];
          *reinterpret_cast<double2*>(tp + m * trow8 + 32 * q) = make_double2(vr, vi);
          if (E1B) {
            er += hq[m][q].x * vr - hq[m][q].y * vi;
            ei += hq[m][q].x * vi + hq[m][q].y * vr;
          }
        }
      }
    }
    tp += 32 * CH;
  }
  };
  if (KS == KSM) chunks(std::true_type{});
  else chunks(std::false_type{});
}

// NMT: 8-row tiles over the occupied orbitals of one spin (ceil(ns/8) <= NMT)
template <int NMT>
__global__ void __launch_bounds__(TH_WARPS * 32, NMT == 1 ? 6 : NMT == 2 ? 4 : NMT == 3 ? 3 : NMT <= 5 ? 2 : 1)
    theta_kernel(ThetaArgs a, int smem_per_warp) {
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
  cplx* colk = A + (size_t)nmax * lda;                                        // [ns] + 3 (slack, zero)
  cplx* prow = colk + nmax + 3;                                               // [2][nc], register path
  int* piv = reinterpret_cast<int*>(prow + (nmax <= 32 ? 4 * ((nmax + 7) / 8 * 8) : 0));  // [nc], behind the pivots [2][nc]
  const int g = lane >> 2, t = lane & 3;
  const int wg = w >> 2, wl = w & 3;

  // 1. O -> shared
  {
    const double2* src = a.OB + ((size_t)w * 2 + s) * a.nsq;
    for (int idx = lane; idx < ns * lda; idx += 32) {
      const double2 v = src[idx];
      A[idx] = {v.x, v.y};
    }
    if (lane < nmax + 3) colk[lane < 3 ? nmax + lane : lane - 3] = {0.0, 0.0};
  }
  __syncwarp();

  // 2. Gauss-Jordan inversion with partial pivoting: in registers up to 32 orbitals per spin, else
  //    in place in shared memory with lanes owning columns j = lane, lane + 32
  cplx sign = {1.0, 0.0};
  double logdet = 0.0;
  constexpr bool GJ_REGS = NMT <= 4;
  if constexpr (GJ_REGS) gj_invert_regs<8 * NMT>(A, lda, ns, prow, piv, lane, sign, logdet);
  for (int k = 0; k < (GJ_REGS ? 0 : ns); ++k) {
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
  for (int k = (GJ_REGS ? 0 : ns) - 1; k >= 0; --k) {
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

  // 3. Theta = Oinv phi^T (DMMA), e1b fused
  double er = 0.0, ei = 0.0;
  {
    const size_t wbase = ((size_t)wg * d.ne + ioff) * d.KC * 32 + wl * 8;
    warp_apply_left<NMT, true, true>(A, lda, ns, d, a.phi + wbase, a.theta + wbase, a.h1rot + (size_t)ioff * d.Mp, lane, er, ei);
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    er += __shfl_xor_sync(0xffffffffu, er, m);
    ei += __shfl_xor_sync(0xffffffffu, ei, m);
  }
  if (lane == 0) a.e1b_part[(size_t)w * 2 + s] = make_double2(er, ei);
}

}  // namespace pxb

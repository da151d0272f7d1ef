// Shared definitions: problem dimensions, HBM layouts, FP64 tensor-core
// (DMMA) and TMA/mbarrier primitives for sm_100a.
//
// HBM layouts ("fragment-major"): every GEMM operand is stored so that the 32
// doubles one warp feeds to one DMMA.8x8x4 (mma.sync.m8n8k4.f64) are
// contiguous (256 B).  Operands then move with plain contiguous copies
// (coalesced LDG.64 per warp, or 1-D TMA bulk copies into shared memory) and
// are read from shared memory without bank conflicts.
//
//   A fragment (8 rows x 4 k):  lane = 4*g + t holds A[g][t]       -> offset lane
//   B fragment (4 k x 8 cols):  lane (g,t) holds B[k=t][n=g]
//   C fragment (8 x 8):         lane (g,t) holds C[g][2t], C[g][2t+1]
//
// For walker matrices (phi, Theta, fields x) the 8 columns of a B/C tile are
// 4 walkers x (re, im): column n = 2*wl + c.  A lane's two C values are then
// the (re, im) of ONE complex number of walker wl = t, which makes the
// exchange-trace and complex epilogues thread-local.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pxb {

struct Dims {
  int M, na, nb, ne, N, W, Wtot;
  int Mp;   // M rounded up to 4
  int KC;   // Mp / 4        k-steps over the basis index
  int M8;   // M rounded up to 8
  int MT;   // M8 / 8        m-tiles over the basis index
  int Wp;   // W rounded up to 4
  int WG;   // Wp / 4        walker groups (n-tiles)
  int Np;   // N rounded up to 8
  int XG;   // Np / 8        m-tiles over the Cholesky index
  int NKC;  // Np / 4        k-steps over the Cholesky index
  int RT;   // ceil(M/2)*KC  active row tiles of the VHS GEMM
  int exp_order;
  int flags;  // PXB_FLAG_* of pxb_config
  double dt, sqrt_dt, ebound, ecore;
};
constexpr int FLAG_FREE_PROJECTION = 1;  // == PXB_FLAG_FREE_PROJECTION
constexpr int FLAG_NO_FORCE_BIAS = 2;    // == PXB_FLAG_NO_FORCE_BIAS
constexpr int FLAG_LOCAL_ENERGY_WEIGHT = 4;  // == PXB_FLAG_LOCAL_ENERGY_WEIGHT
constexpr int FLAG_COMPLEX_ONE_BODY = 8;     // == PXB_FLAG_COMPLEX_ONE_BODY
constexpr int FLAG_COMPLEX_CHOLESKY = 16;    // == PXB_FLAG_COMPLEX_CHOLESKY
#ifndef PXB_MAX_DETS
#define PXB_MAX_DETS 8  // == include/pauxy_b200.h
#endif

__host__ __device__ inline int round_up(int x, int m) { return (x + m - 1) / m * m; }
// Complex Cholesky vectors / trial orbitals: every GEMM with a constant A operand runs as the REAL
// fragment-major GEMM over a doubled k range, [Re A | Im A] [B ; i B]; kmul is that factor.
__host__ __device__ inline int kmul(const Dims& d) { return (d.flags & FLAG_COMPLEX_CHOLESKY) ? 2 : 1; }

// ---- layout index helpers (doubles) -----------------------------------------
// OF: orbital-fragment layout of phi / Theta: [WG][ne][KC][wl 4][t 4][c 2]
__host__ __device__ inline size_t of_index(const Dims& d, int w, int i, int p, int c) {
  return ((((size_t)(w >> 2) * d.ne + i) * d.KC + (p >> 2)) * 4 + (w & 3)) * 8 + (p & 3) * 2 + c;
}
__host__ __device__ inline size_t of_size(const Dims& d) { return (size_t)d.WG * d.ne * d.KC * 32; }
// XF: field layout [WG][NKC][wl 4][t 4][c 2]
__host__ __device__ inline size_t xf_index(const Dims& d, int w, int n, int c) {
  return (((size_t)(w >> 2) * d.NKC + (n >> 2)) * 4 + (w & 3)) * 8 + (n & 3) * 2 + c;
}
__host__ __device__ inline size_t xf_size(const Dims& d) { return (size_t)d.WG * d.NKC * 32; }
// RF: half-rotated Cholesky, per spin [XG][n_s][kmul KC][g 8][t 4]; spin 1 follows spin 0
// (complex: k-steps [0, KC) of an orbital hold the real part, [KC, 2 KC) the imaginary part)
__host__ __device__ inline size_t rf_spin_base(const Dims& d, int s) {
  return s == 0 ? 0 : (size_t)d.XG * d.na * d.KC * 32 * kmul(d);
}
__host__ __device__ inline size_t rf_size(const Dims& d) { return (size_t)d.XG * d.ne * d.KC * 32 * kmul(d); }
// LF: Cholesky for the VHS GEMM [RT][NKC][g' 8][t 4]; row tile rt = (mt*4+s)*KC + kc
// holds rows (p, q) = (8mt + 2s + (g'>>2), 4kc + (g'&3))
// (complex: [RT][2 NKC]..., real part in k-steps [0, NKC), imaginary part in [NKC, 2 NKC))
__host__ __device__ inline size_t lf_size(const Dims& d) { return (size_t)d.RT * d.NKC * 32 * kmul(d); }
// VF: VHS per walker as Taylor A-operand [W][MT][KC][c 2][g 8][t 4]
__host__ __device__ inline size_t vf_walker(const Dims& d) { return (size_t)d.MT * d.KC * 64; }
// BF: one-body propagator [2][MT][KC][g 8][t 4]
__host__ __device__ inline size_t bf_size(const Dims& d) { return (size_t)2 * d.MT * d.KC * 32; }

// lane offset of a B fragment element inside its 32-double block
__device__ __forceinline__ int b_lane_offset(int lane) {
  int g = lane >> 2, t = lane & 3;
  return (g >> 1) * 8 + t * 2 + (g & 1);
}

// ---- DMMA ------------------------------------------------------------------
// mma.sync.m8n8k4.f64 == one SASS DMMA.8x8x4 on sm_100a (the wider f64 shapes
// m16n8k4/k8/k16 are split into DMMA.8x8x4 by ptxas, checked with cuobjdump).
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
      : "+d"(c0), "+d"(c1)
      : "d"(a), "d"(b));
}

__device__ __forceinline__ double ldg_nc(const double* p) { return __ldg(p); }
__device__ __forceinline__ double2 ldg_nc2(const double* p) { return __ldg(reinterpret_cast<const double2*>(p)); }

// ---- mbarrier + TMA bulk copy (cp.async.bulk, SASS UBLKCP) ---------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, unsigned bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// Consumer side of a TMA ring: a warp hands a stage back to the producer after its last read of it.
// Product form: the warp's reads are ordered before lane 0's (release) arrive by __syncwarp().
// -DPXB_RACECHECK_STRICT (a checker build, tools/gpu.sh sanitize): every lane arrives itself, which
// is what compute-sanitizer's racecheck can follow (it does not carry the __syncwarp ordering over
// to another lane's mbarrier arrive); used to show that the hazards it reports are of that kind.
#ifdef PXB_RACECHECK_STRICT
constexpr unsigned kReleaseArrivals = 32;
__device__ __forceinline__ void ring_release(uint64_t* bar, int) { mbar_arrive(bar); }
#else
constexpr unsigned kReleaseArrivals = 1;
__device__ __forceinline__ void ring_release(uint64_t* bar, int lane) {
  __syncwarp();
  if (lane == 0) mbar_arrive(bar);
}
#endif

// 1-D bulk copy global -> shared, completion signalled on an mbarrier.
// bytes must be a multiple of 16; both addresses 16-byte aligned.
__device__ __forceinline__ void tma_bulk_g2s(void* smem_dst, const void* gmem_src, unsigned bytes,
                                             uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(smem_u32(smem_dst)),
      "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}

// ---- complex helpers ----------------------------------------------------------
struct cplx {
  double re, im;
};
__host__ __device__ inline cplx cmul(cplx a, cplx b) {
  return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
__host__ __device__ inline cplx cadd(cplx a, cplx b) { return {a.re + b.re, a.im + b.im}; }
__host__ __device__ inline cplx csub(cplx a, cplx b) { return {a.re - b.re, a.im - b.im}; }
__host__ __device__ inline cplx cdiv(cplx a, cplx b) {
  // Smith's algorithm (as numpy / C99 do)
  if (fabs(b.re) >= fabs(b.im)) {
    double r = b.im / b.re, den = b.re + b.im * r;
    return {(a.re + a.im * r) / den, (a.im - a.re * r) / den};
  } else {
    double r = b.re / b.im, den = b.re * r + b.im;
    return {(a.re * r + a.im) / den, (a.im * r - a.re) / den};
  }
}

}  // namespace pxb

// Taylor expansion  phi <- sum_{n<=order} VHS^n / n! phi   per walker
// (propagation/continuous.py:82-111; both spin blocks use the same VHS,
// :169-171), evaluated in Horner form
//     S_order = phi,   S_{n-1} = phi + (VHS S_n) / n,   result = S_0
// which is the same polynomial as the reference's running sum (Temp = VHS Temp / n;
// phi += Temp) with rounding differences at the 1e-16 level, and needs no
// read-modify-write of phi between orders.
//
// complex x complex as two real DMMA streams:
//     C = A_re * B^ + A_im * (i B)^        B^ = [.. (o,re) (o,im) ..] real columns
// One CTA = one walker x one chunk of orbitals (columns are independent).
// S_n lives in shared memory in B-fragment order, the accumulators of the next
// iterate in registers; VHS (A-fragment order, written by the VHS GEMM
// epilogue) streams from L2 once per order.
#pragma once
#include "pxb_common.cuh"

namespace pxb {

struct TaylorArgs {
  const double* VF;  // [W][MT][KC][2][32]
  double* phi;       // OF layout, updated in place
  const int* active; // optional
  Dims d;
  int ochunk;        // orbitals per CTA (multiple of 4)
  int nchunks;
};

// position of B element (k = t, n = g) inside a 32-double block: the 16 lanes of a
// half warp (g < 4 or g >= 4) read 16 consecutive doubles -> no bank conflicts
__device__ __forceinline__ int tb_off(int k, int n) { return (n >> 2) * 16 + k * 4 + (n & 3); }

// WMT m-tiles per warp, NT n-tiles per CTA (compile time: no predication around the DMMAs;
// orbitals beyond the chunk are zero columns), NWARPS warps per CTA
template <int WMT, int NT, int NWARPS, int MINB>
__global__ void __launch_bounds__(NWARPS * 32, MINB) taylor_kernel(TaylorArgs a) {
  extern __shared__ __align__(16) double Ts[];  // [KC][NT][32]
  const Dims& d = a.d;
  const int w = blockIdx.x / a.nchunks, chunk = blockIdx.x % a.nchunks;
  if (a.active != nullptr && a.active[w] == 0) return;
  const int o0 = chunk * a.ochunk;
  const int no = min(a.ochunk, d.ne - o0);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int g = lane >> 2, t = lane & 3;
  const int wg = w >> 2, wl = w & 3;

  // S_order = phi columns o0..o0+no (zero padded to 4*NT orbitals)
  for (int idx = tid; idx < d.KC * NT * 16; idx += blockDim.x) {
    // idx -> (kc, nt, tt, oo): (re, im) of orbital o0 + 4nt + oo at basis index p = 4kc + tt
    const int oo = idx & 3, tt = (idx >> 2) & 3, r = idx >> 4;
    const int nt = r % NT, kc = r / NT;
    const int ol = 4 * nt + oo;
    double2 v = make_double2(0.0, 0.0);
    if (ol < no)
      v = *reinterpret_cast<const double2*>(a.phi + (((size_t)wg * d.ne + o0 + ol) * d.KC + kc) * 32 +
                                            wl * 8 + tt * 2);
    *reinterpret_cast<double2*>(Ts + ((size_t)kc * NT + nt) * 32 + tb_off(tt, 2 * oo)) = v;
  }
  __syncthreads();

  const int mt0 = warp * WMT;
  const double* Aw = a.VF + (size_t)w * vf_walker(d);
  const bool warp_active = mt0 < d.MT;
  const int boff = tb_off(t, g), boffp = tb_off(t, g ^ 1);
  const double sgn = (g & 1) ? 1.0 : -1.0;  // (i B)^: (re, im) -> (-im, re)

  for (int n = d.exp_order; n >= 1; --n) {
    double acc[WMT][NT][2];
#pragma unroll
    for (int i = 0; i < WMT; ++i)
#pragma unroll
      for (int j = 0; j < NT; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;

    if (warp_active) {
      const double* Ap[WMT];
#pragma unroll
      for (int i = 0; i < WMT; ++i) Ap[i] = Aw + (size_t)min(mt0 + i, d.MT - 1) * d.KC * 64 + lane;
      double ar[WMT], ai[WMT], nr[WMT], ni[WMT];
#pragma unroll
      for (int i = 0; i < WMT; ++i) {
        ar[i] = ldg_nc(Ap[i]);
        ai[i] = ldg_nc(Ap[i] + 32);
      }
      for (int kc = 0; kc < d.KC; ++kc) {
        const int kn = (kc + 1 < d.KC ? kc + 1 : kc) * 64;
#pragma unroll
        for (int i = 0; i < WMT; ++i) {
          nr[i] = ldg_nc(Ap[i] + kn);
          ni[i] = ldg_nc(Ap[i] + kn + 32);
        }
        const double* Tk = Ts + (size_t)kc * NT * 32;
#pragma unroll
        for (int j = 0; j < NT; ++j) {
          const double b = Tk[j * 32 + boff];
          const double bq = sgn * Tk[j * 32 + boffp];
#pragma unroll
          for (int i = 0; i < WMT; ++i) {
            dmma(acc[i][j][0], acc[i][j][1], ar[i], b);
            dmma(acc[i][j][0], acc[i][j][1], ai[i], bq);
          }
        }
#pragma unroll
        for (int i = 0; i < WMT; ++i) {
          ar[i] = nr[i];
          ai[i] = ni[i];
        }
      }
    }
    __syncthreads();  // every warp has finished reading S_n
    if (warp_active) {
#pragma unroll
      for (int i = 0; i < WMT; ++i) {
        const int mt = mt0 + i;
        const int p = 8 * mt + g;
        const int kc2 = p >> 2, t2 = p & 3;
        if (mt < d.MT && kc2 < d.KC) {
#pragma unroll
          for (int j = 0; j < NT; ++j) {
            const int ol = 4 * j + t;
            double2* gp = reinterpret_cast<double2*>(
                a.phi + (((size_t)wg * d.ne + o0 + min(ol, no - 1)) * d.KC + kc2) * 32 + wl * 8 + t2 * 2);
            double2 p0 = make_double2(0.0, 0.0);
            if (ol < no) p0 = *gp;
            // the reference divides (Temp = VHS.dot(Temp) / n); so do we
            const double vr = p0.x + acc[i][j][0] / (double)n;
            const double vi = p0.y + acc[i][j][1] / (double)n;
            if (n > 1) {
              *reinterpret_cast<double2*>(Ts + ((size_t)kc2 * NT + j) * 32 + tb_off(t2, 2 * t)) =
                  make_double2(vr, vi);
            } else if (ol < no) {
              *gp = make_double2(vr, vi);
            }
          }
        }
      }
    }
    __syncthreads();
  }
}

}  // namespace pxb

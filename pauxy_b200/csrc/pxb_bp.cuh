// Back propagation for the generic Hamiltonian:
//   walkers/stack.py:5-127          FieldConfig: the shifted fields x of every step, per walker
//   propagation/generic.py:180-213  B(c) = BH1 exp6(i sqrt(dt) L.c) BH1
//   propagation/generic.py:253-290  back_propagate_generic: phi_bp <- B(c)^dagger phi_bp, reversed
//   estimators/back_propagation.py:127-225  G = gab(phi_bp, phi_old)^T, sum_w weight_w G_w
//
// With real symmetric Cholesky matrices (checked at set-up) and a real symmetric BH1,
//   B(c)^dagger = BH1 exp6(i sqrt(dt) L.(-conj c)) BH1,
// so one back-propagation step is the forward propagation chain (VHS GEMM, one-body GEMM,
// Taylor kernel, one-body GEMM) on the trial determinant with the stored field negated and
// conjugated.  Only the pieces that have no forward counterpart live here.
#pragma once
#include "pxb_common.cuh"

namespace pxb {

// FC: field history [WG][nbp * NKC][4 walkers][4 fields][re, im]: per walker group one row of
// NKC fragments per stored step, i.e. step s of all walkers is a strided copy of the XF buffer.
__host__ __device__ inline size_t fc_rows(const Dims& d, int nbp) { return (size_t)nbp * d.NKC; }
__host__ __device__ inline size_t fc_size(const Dims& d, int nbp) { return (size_t)d.WG * fc_rows(d, nbp) * 32; }

// XF <- -conj(FC[step])
__global__ void bp_field_kernel(const double* __restrict__ FC, double* __restrict__ XF, Dims d, int nbp,
                                int step) {
  const size_t per_wg = (size_t)d.NKC * 16;  // double2 elements of one walker group and step
  const size_t total = (size_t)d.WG * per_wg;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const size_t wg = idx / per_wg, r = idx % per_wg;
    const double2 v = reinterpret_cast<const double2*>(FC)[(wg * nbp + step) * per_wg + r];
    reinterpret_cast<double2*>(XF)[idx] = make_double2(-v.x, v.y);
  }
}

// O[(w, s)][i][j] = sum_p phi_old[p, i] conj(phi_bp[p, j])  (the overlap of gab(), transposed the
// way theta_kernel expects it: Theta = O^-1 phi_old^T).  One CTA per (walker, spin).
__global__ void __launch_bounds__(128) bp_overlap_kernel(const double* __restrict__ phi_old,
                                                         const double* __restrict__ phi_bp,
                                                         double2* __restrict__ OB, Dims d, int nld, int nsq) {
  const int w = blockIdx.x >> 1, s = blockIdx.x & 1;
  if (w >= d.Wp) return;
  const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
  double2* out = OB + ((size_t)w * 2 + s) * nsq;
  for (int e = threadIdx.x; e < ns * ns; e += blockDim.x) {
    const int i = e / ns, j = e % ns;
    const double2* a = reinterpret_cast<const double2*>(phi_old + of_index(d, w, ioff + i, 0, 0));
    const double2* b = reinterpret_cast<const double2*>(phi_bp + of_index(d, w, ioff + j, 0, 0));
    double re = 0.0, im = 0.0;
    for (int kc = 0; kc < d.KC; ++kc) {
#pragma unroll
      for (int t = 0; t < 4; ++t) {
        const double2 x = a[(size_t)kc * 16 + t], y = b[(size_t)kc * 16 + t];  // 4 walkers x 4 p per kc
        re += x.x * y.x + x.y * y.y;   // x * conj(y)
        im += x.y * y.x - x.x * y.y;
      }
    }
    out[(size_t)i * nld + j] = make_double2(re, im);
  }
}

// partial[chunk][s][p][q] = sum_{w in chunk} weight_w sum_i conj(phi_bp_w[p, i]) Theta_w[i, q]
// 64 x 64 output tile per CTA, 4 x 4 complex accumulators per thread, K = (walker, orbital) pairs
// staged 8 at a time through shared memory.
constexpr int BPR_T = 64, BPR_K = 8;
struct BpRdmArgs {
  const double* phi_bp;   // OF
  const double* theta;    // OF, Theta_bp = (phi_old^T conj(phi_bp))^-1 phi_old^T
  const double2* weight;  // [W] complex estimator weights (bp_weight_kernel)
  double2* part;          // [nchunks][2][M][M]
  Dims d;
  int nchunks, wchunk, tiles;
};

__global__ void __launch_bounds__(256) bp_rdm_kernel(BpRdmArgs a) {
  __shared__ double2 As[BPR_K][BPR_T];
  __shared__ double2 Ts[BPR_K][BPR_T];
  const Dims& d = a.d;
  int b = blockIdx.x;
  const int tq = b % a.tiles;
  b /= a.tiles;
  const int tp = b % a.tiles;
  b /= a.tiles;
  const int s = b & 1, chunk = b >> 1;
  const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
  const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
  const int p0 = tp * BPR_T, q0 = tq * BPR_T;
  double acc[4][4][2];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int w0 = chunk * a.wchunk, w1 = min(w0 + a.wchunk, d.W);
  const long long ktot = (long long)max(w1 - w0, 0) * ns;
  for (long long k0 = 0; k0 < ktot; k0 += BPR_K) {
    // stage: 2 x (8 x 64) complex, 256 threads -> 2 elements of each matrix per thread
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int e = threadIdx.x + 256 * r;
      const int kk = e >> 6, c = e & 63;
      const long long k = k0 + kk;
      double2 va = make_double2(0.0, 0.0), vt = va;
      if (k < ktot) {
        const int w = w0 + (int)(k / ns), i = ioff + (int)(k % ns);
        const double2 wt = a.weight[w];
        if (wt.x != 0.0 || wt.y != 0.0) {
          if (p0 + c < d.M) {
            const double2 x = *reinterpret_cast<const double2*>(a.phi_bp + of_index(d, w, i, p0 + c, 0));
            va = make_double2(wt.x * x.x + wt.y * x.y, wt.y * x.x - wt.x * x.y);  // weight * conj(phi_bp)
          }
          if (q0 + c < d.M) vt = *reinterpret_cast<const double2*>(a.theta + of_index(d, w, i, q0 + c, 0));
        }
      }
      As[kk][c] = va;
      Ts[kk][c] = vt;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BPR_K; ++kk) {
      double2 x[4], y[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) x[i] = As[kk][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) y[j] = Ts[kk][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          acc[i][j][0] += x[i].x * y[j].x - x[i].y * y[j].y;
          acc[i][j][1] += x[i].x * y[j].y + x[i].y * y[j].x;
        }
    }
    __syncthreads();
  }
  double2* out = a.part + ((size_t)chunk * 2 + s) * d.M * d.M;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int p = p0 + ty * 4 + i;
    if (p >= d.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int q = q0 + tx * 4 + j;
      if (q < d.M) out[(size_t)p * d.M + q] = make_double2(acc[i][j][0], acc[i][j][1]);
    }
  }
}

// Estimator weight of back_propagation.py:187-196: BP-PhL walker.weight (mode 0), BP-PRes
// weight * prod(I/|I|) (mode 1, restore_weights = "partial"), BP-Pres weight * prod(I/|I|) /
// prod(cosine_fac) (mode 2, "full")
__global__ void bp_weight_kernel(const double* __restrict__ weight, const double2* __restrict__ bpfac,
                                 double2* __restrict__ out, int mode, int W) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  double2 f = make_double2(1.0, 0.0);
  if (mode >= 1) f = bpfac[2 * w];
  if (mode == 2) {
    const double c = bpfac[2 * w + 1].x;
    f = make_double2(f.x / c, f.y / c);
  }
  const double wt = weight[w];
  out[w] = make_double2(wt * f.x, wt * f.y);
}

// rdm[s][p][q] += sum_chunk part (fixed order); denom += sum_w weight (fixed tree), thread 0 of CTA 0
__global__ void __launch_bounds__(256) bp_reduce_kernel(const double2* __restrict__ part, double2* __restrict__ rdm,
                                                        double2* __restrict__ denom,
                                                        const double2* __restrict__ weight, Dims d, int nchunks) {
  const size_t n = (size_t)2 * d.M * d.M;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n;
       idx += (size_t)gridDim.x * blockDim.x) {
    double re = 0.0, im = 0.0;
    for (int c = 0; c < nchunks; ++c) {
      const double2 v = part[(size_t)c * n + idx];
      re += v.x;
      im += v.y;
    }
    rdm[idx].x += re;
    rdm[idx].y += im;
  }
  if (blockIdx.x == 0) {
    __shared__ double2 red[256];
    double sr = 0.0, si = 0.0;
    for (int w = threadIdx.x; w < d.W; w += 256) {
      sr += weight[w].x;
      si += weight[w].y;
    }
    red[threadIdx.x] = make_double2(sr, si);
    __syncthreads();
    for (int m = 128; m > 0; m >>= 1) {
      if (threadIdx.x < m) {
        red[threadIdx.x].x += red[threadIdx.x + m].x;
        red[threadIdx.x].y += red[threadIdx.x + m].y;
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      denom[0].x += red[0].x;
      denom[0].y += red[0].y;
    }
  }
}

}  // namespace pxb

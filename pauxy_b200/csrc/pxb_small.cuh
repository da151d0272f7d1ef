// Memory/latency-bound stages: layout packing, overlap + LU inverse +
// rotated Green's function, modified Gram-Schmidt re-orthogonalisation, field
// shift, weight update, energy assembly, estimator accumulation, comb.
#pragma once
#include "pxb_common.cuh"

namespace pxb {

// ============================================================================
// packing of the constant operands (setup)
// ============================================================================
// flag[0] |= 1 if a supposedly real input has a non-zero imaginary part
// rt_map (optional): compact list of the row tiles that are kept (upper triangle of a symmetric L)
// hs_pot: real [M*M][N] doubles, or (FLAG_COMPLEX_CHOLESKY) complex128: then the k range is doubled
__global__ void pack_lf_kernel(const double* __restrict__ hs_pot, double* __restrict__ LF, Dims d,
                               const int* __restrict__ rt_map, int nrt) {
  const int km = kmul(d);
  const size_t total = (size_t)nrt * d.NKC * 32 * km;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int tt = idx & 3, gp = (idx >> 2) & 7;
    const size_t r = idx >> 5;
    const int kc2 = r % (d.NKC * km);
    const int kcn = kc2 < d.NKC ? kc2 : kc2 - d.NKC, part = kc2 < d.NKC ? 0 : 1;
    const int rtc = r / (d.NKC * km);
    const int rt = rt_map != nullptr ? rt_map[rtc] : rtc;
    const int kc = rt % d.KC, ms = rt / d.KC, s = ms & 3, mtv = ms >> 2;
    const int p = 8 * mtv + 2 * s + (gp >> 2), q = 4 * kc + (gp & 3), n = 4 * kcn + tt;
    double v = 0.0;
    if (p < d.M && q < d.M && n < d.N) v = hs_pot[(((size_t)p * d.M + q) * d.N + n) * km + part];
    LF[idx] = v;
  }
}

// flag |= 16 unless hs_pot[(p,q), n] == hs_pot[(q,p), n] for all p, q, n (then VHS is symmetric and
// only its upper triangle is computed by the GEMM)
__global__ void hs_symmetry_kernel(const double* __restrict__ hs_pot, Dims d, int* flag) {
  const size_t total = (size_t)d.M * d.M * d.N;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int n = idx % d.N;
    const size_t pq = idx / d.N;
    const int q = pq % d.M, p = pq / d.M;
    if (p < q) {
      const int km = kmul(d);
      for (int part = 0; part < km; ++part) {
      const double a = hs_pot[idx * km + part], b = hs_pot[(((size_t)q * d.M + p) * d.N + n) * km + part];
      if (a != b) {
        // 16: not bit-symmetric; 32: not symmetric even to rounding (back propagation then needs L^T)
        atomicOr(flag, fabs(a - b) > 1e-12 * (fabs(a) + fabs(b)) + 1e-14 ? 48 : 16);
      }
      }
    }
  }
}

__global__ void pack_rf_kernel(const double2* __restrict__ rchol, double* __restrict__ RF, Dims d,
                               int* flag) {
  const size_t total = rf_size(d);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    int s = 0;
    size_t rel = idx;
    const size_t b1 = rf_spin_base(d, 1);
    if (idx >= b1) {
      s = 1;
      rel = idx - b1;
    }
    const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
    const int tt = rel & 3, g = (rel >> 2) & 7;
    size_t r = rel >> 5;
    const int km = kmul(d);
    const int pc2 = r % (d.KC * km);
    const int pc = pc2 < d.KC ? pc2 : pc2 - d.KC;
    r /= d.KC * km;
    const int il = r % ns, xg = r / ns;
    const int p = 4 * pc + tt, n = 8 * xg + g;
    double v = 0.0;
    if (p < d.M && n < d.N) {
      double2 z = rchol[((size_t)(ioff + il) * d.M + p) * d.N + n];
      v = pc2 < d.KC ? z.x : z.y;
      if (km == 1 && z.y != 0.0) atomicOr(flag, 1);
    }
    RF[idx] = v;
  }
}

__global__ void pack_bf_kernel(const double2* __restrict__ bh1, double* __restrict__ BF, Dims d,
                               int* flag, int transpose) {
  const size_t total = bf_size(d);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int tt = idx & 3, g = (idx >> 2) & 7;
    size_t r = idx >> 5;
    const int kc = r % d.KC;
    r /= d.KC;
    const int mt = r % d.MT, s = r / d.MT;
    const int p = 8 * mt + g, q = 4 * kc + tt;
    double v = 0.0;
    if (p < d.M && q < d.M) {
      double2 z = transpose ? bh1[((size_t)s * d.M + q) * d.M + p] : bh1[((size_t)s * d.M + p) * d.M + q];
      v = z.x;
      if (z.y != 0.0) atomicOr(flag, 2);
    }
    BF[idx] = v;
  }
}

// Complex one-body propagator (multi-determinant trials with complex CI coefficients make the
// mean-field shift, hence BH1, genuinely complex): BH1 phi = [Re BH1 | Im BH1] [phi ; i phi], i.e.
// the REAL fragment-major GEMM with the k range doubled.  BF2[s][mt][kc'][g][t]: kc' < KC the real
// part of BH1[s][8mt+g][4kc'+t], kc' >= KC the imaginary part.
__global__ void pack_bf2_kernel(const double2* __restrict__ bh1, double* __restrict__ BF2, Dims d) {
  const size_t total = 2 * bf_size(d);
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int tt = idx & 3, g = (idx >> 2) & 7;
    size_t r = idx >> 5;
    const int kc2 = r % (2 * d.KC);
    r /= 2 * d.KC;
    const int mt = r % d.MT, s = r / d.MT;
    const int kc = kc2 < d.KC ? kc2 : kc2 - d.KC;
    const int p = 8 * mt + g, q = 4 * kc + tt;
    double v = 0.0;
    if (p < d.M && q < d.M) {
      const double2 z = bh1[((size_t)s * d.M + p) * d.M + q];
      v = kc2 < d.KC ? z.x : z.y;
    }
    BF2[idx] = v;
  }
}
// [B ; i B] stacked along k for rows of kc 32-double blocks (OF rows: kc = KC, nrows = WG ne; XF rows:
// kc = NKC, nrows = WG): out[row][2 kc][wl][t][c]
__global__ void stack_rows_kernel(const double* __restrict__ in, double* __restrict__ out, size_t nrows, int kc) {
  const size_t total = nrows * kc * 16;  // complex elements
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int e = idx & 15;
    const size_t r = idx >> 4;
    const int k = r % kc;
    const size_t row = r / kc;
    const double2 v = *reinterpret_cast<const double2*>(in + (row * kc + k) * 32 + e * 2);
    double* o = out + (row * 2 * kc + k) * 32 + e * 2;
    *reinterpret_cast<double2*>(o) = v;
    *reinterpret_cast<double2*>(o + (size_t)kc * 32) = make_double2(-v.y, v.x);
  }
}
// psiT[p][j] (real, [Mp][ne]), h1rot [ne][Mp] complex, vbar [Np] complex
__global__ void pack_small_kernel(const double2* __restrict__ psi, const double2* __restrict__ h1rot,
                                  const double2* __restrict__ mf, double* __restrict__ psiT,
                                  double2* __restrict__ h1r, double2* __restrict__ vbar, Dims d,
                                  int* flag) {
  const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
  for (int idx = tid; idx < d.ne * d.Mp; idx += nth) {
    const int j = idx / d.Mp, p = idx % d.Mp;
    double v = 0.0;
    double2 h = make_double2(0.0, 0.0);
    if (p < d.M) {
      double2 z = psi[(size_t)p * d.ne + j];
      v = z.x;
      if (z.y != 0.0) atomicOr(flag, 4);
      h = h1rot[(size_t)j * d.M + p];
    }
    psiT[(size_t)p * d.ne + j] = v;
    h1r[idx] = h;
  }
  if (mf != nullptr)
    for (int n = tid; n < d.Np; n += nth) vbar[n] = n < d.N ? mf[n] : make_double2(0.0, 0.0);
}

// psi as DMMA B-fragments for the overlap GEMM: PF[s][jt][pc][lane=(g,t)] = psi[4pc+t][ioff_s + 8jt+g]
// (complex orbitals: the operand is conj(psi); k-steps [KC, 2 KC) hold its imaginary part, -Im psi)
__global__ void pack_pf_kernel(const double2* __restrict__ psi, double* __restrict__ PF, Dims d) {
  const int jt0 = (d.na + 7) >> 3, jt1 = (d.nb + 7) >> 3;
  const int km = kmul(d);
  const size_t total = (size_t)(jt0 + jt1) * d.KC * 32 * km;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int lane = idx & 31, g = lane >> 2, t = lane & 3;
    size_t r = idx >> 5;
    const int pc2 = r % (d.KC * km);
    const int pc = pc2 < d.KC ? pc2 : pc2 - d.KC;
    int jt = r / (d.KC * km), s = 0;
    if (jt >= jt0) {
      s = 1;
      jt -= jt0;
    }
    const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
    const int p = 4 * pc + t, j = 8 * jt + g;
    const double2 z = (p < d.M && j < ns) ? psi[(size_t)p * d.ne + ioff + j] : make_double2(0.0, 0.0);
    PF[idx] = pc2 < d.KC ? z.x : -z.y;
  }
}

// ============================================================================
// walker matrix packing: reference layout [W][M][ne] complex <-> OF
// ============================================================================
__global__ void phi_to_of_kernel(const double2* __restrict__ phi, double* __restrict__ of, Dims d,
                                 int bcast_single) {
  // one thread per (w, i, p) complex element of the padded OF array
  const size_t total = (size_t)d.Wp * d.ne * d.Mp;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int p = idx % d.Mp;
    size_t r = idx / d.Mp;
    const int i = r % d.ne, w = r / d.ne;
    double2 v = make_double2(0.0, 0.0);
    if (p < d.M) {
      if (bcast_single)
        v = phi[(size_t)p * d.ne + i];
      else if (w < d.W)
        v = phi[((size_t)w * d.M + p) * d.ne + i];
      else
        v = phi[((size_t)0 * d.M + p) * d.ne + i];  // padding walkers mirror walker 0
    }
    *reinterpret_cast<double2*>(of + of_index(d, w, i, p, 0)) = v;
  }
}

// generic unpack of an OF-layout array to [W][rows=ne][M] (transpose=0, Theta) or
// [W][M][ne] (transpose=1, phi)
__global__ void of_to_natural_kernel(const double* __restrict__ of, double2* __restrict__ out, Dims d,
                                     int as_phi) {
  const size_t total = (size_t)d.W * d.ne * d.M;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int p = idx % d.M;
    size_t r = idx / d.M;
    const int i = r % d.ne, w = r / d.ne;
    double2 v = *reinterpret_cast<const double2*>(of + of_index(d, w, i, p, 0));
    if (as_phi)
      out[((size_t)w * d.M + p) * d.ne + i] = v;
    else
      out[((size_t)w * d.ne + i) * d.M + p] = v;
  }
}

__global__ void vf_to_natural_kernel(const double* __restrict__ VF, double2* __restrict__ out, Dims d) {
  const size_t total = (size_t)d.W * d.M * d.M;
  for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (size_t)gridDim.x * blockDim.x) {
    const int q = idx % d.M;
    size_t r = idx / d.M;
    const int p = r % d.M, w = r / d.M;
    const double* b = VF + (size_t)w * vf_walker(d) + ((size_t)(p >> 3) * d.KC + (q >> 2)) * 64 +
                      (p & 7) * 4 + (q & 3);
    out[idx] = make_double2(b[0], b[32]);
  }
}

// ============================================================================
// Philox4x32-10 + Box-Muller (throughput mode fields)
// ============================================================================
__device__ __forceinline__ void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c[0]), lo0 = 0xD2511F53u * c[0];
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c[2]), lo1 = 0xCD9E8D57u * c[2];
    const uint32_t n0 = hi1 ^ c[1] ^ k0, n2 = hi0 ^ c[3] ^ k1;
    c[0] = n0;
    c[1] = lo1;
    c[2] = n2;
    c[3] = lo0;
    k0 += 0x9E3779B9u;
    k1 += 0xBB67AE85u;
  }
}
// two independent N(0,1) from one counter (walker, field pair, step)
__device__ __forceinline__ void philox_normal2(uint64_t seed, uint64_t step, uint64_t gw, uint32_t pair,
                                               double& z0, double& z1) {
  uint32_t c[4] = {(uint32_t)gw, (uint32_t)(gw >> 32) ^ (pair * 0x85ebca6bu), pair, (uint32_t)step};
  philox4x32_10(c, (uint32_t)seed ^ (uint32_t)(step >> 32), (uint32_t)(seed >> 32));
  const double u0 = ((double)c[0] * 4294967296.0 + (double)c[1] + 0.5) * (1.0 / 18446744073709551616.0);
  const double u1 = ((double)c[2] * 4294967296.0 + (double)c[3] + 0.5) * (1.0 / 18446744073709551616.0);
  const double r = sqrt(-2.0 * log(u0));
  double sn, cs;
  sincospi(2.0 * u1, &sn, &cs);
  z0 = r * cs;
  z1 = r * sn;
}

// ============================================================================
// Per-step scalars.  They live in device memory (not in kernel arguments) so that the launch
// sequence of a whole driver step can be captured once as a CUDA graph and replayed with new
// values (pxb_step): one tiny setter launch in front of the graph instead of re-parametrising
// its kernel nodes.
// ============================================================================
struct StepParams {
  unsigned long long seed, step;  // Philox key / counter, driver step (weight cap needs step > 1)
  long long walker_offset;        // global index of this device's first walker
  double eshift;                  // energy shift of the block (continuous.py:202-214)
  double eshift_im;               // its imaginary part (zero in the driver; the reference's own
                                  // propagation tests pass a complex trial energy)
  double comb_r;                  // the comb's uniform draw (handler.py:275)
};
__global__ void set_step_params_kernel(StepParams* dst, StepParams v, int mask) {
  if (mask & 1) {
    dst->seed = v.seed;
    dst->step = v.step;
    dst->walker_offset = v.walker_offset;
    dst->eshift = v.eshift;
    dst->eshift_im = v.eshift_im;
  }
  if (mask & 2) dst->comb_r = v.comb_r;
}

// ============================================================================
// K2b: field shift (propagation/continuous.py:133-158, generic.py:152)
//   xbar = -sqrt(dt) (i V - vbar), clip |xbar| > 1, x = xi - xbar,
//   cmf = -sqrt(dt) x.vbar, cfb = xi.xbar - xbar.xbar/2.   One warp per walker.
// ============================================================================
struct FieldArgs {
  const double2* X;      // [2][Wp][Np]
  const double* xi;      // [W][N] or null (Philox)
  const double2* vbar;   // [Np]
  const int* active;
  double* XF;            // field fragment layout out
  double2* xbar_out;     // [Wp][Np]
  double2* xs_out;       // [Wp][Np] natural-layout copy of x
  double2* cmfcfb;       // [Wp][2]
  long long* counters;
  Dims d;
  const StepParams* sp;  // seed, step, walker_offset
};

__global__ void __launch_bounds__(256) field_kernel(FieldArgs a) {
  const Dims& d = a.d;
  const uint64_t seed = a.sp->seed, step = a.sp->step;
  const long long walker_offset = a.sp->walker_offset;
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= d.Wp) return;
  const bool live = (w < d.W) && (a.active == nullptr || a.active[w] != 0);
  const double2* Xa = a.X + (size_t)w * d.Np;
  const double2* Xb = a.X + ((size_t)d.Wp + w) * d.Np;
  double cmf_r = 0, cmf_i = 0, s1_r = 0, s1_i = 0, s2_r = 0, s2_i = 0;
  int ntrig = 0;
  // lanes walk pairs of fields so that a Philox call yields both normals of a pair
  for (int n0 = 2 * lane; n0 < d.Np; n0 += 64) {
    double z[2] = {0.0, 0.0};
    if (live) {
      if (a.xi != nullptr) {
        if (n0 < d.N) z[0] = a.xi[(size_t)w * d.N + n0];
        if (n0 + 1 < d.N) z[1] = a.xi[(size_t)w * d.N + n0 + 1];
      } else {
        philox_normal2(seed, step, (uint64_t)(walker_offset + w), (uint32_t)(n0 >> 1), z[0], z[1]);
      }
    }
#pragma unroll
    for (int u = 0; u < 2; ++u) {
      const int n = n0 + u;
      double xr = 0.0, xi_ = 0.0, br = 0.0, bi = 0.0;
      if (live && n < d.N) {
        const double2 mf = a.vbar[n];
        if (!(d.flags & FLAG_NO_FORCE_BIAS)) {   // continuous.py:136-138: xbar = 0 otherwise
          const double2 va = Xa[n], vb = Xb[n];
          const double vr = va.x + vb.x, vi = va.y + vb.y;
          // xbar = -sqrt_dt * (1j*V - vbar)
          br = -d.sqrt_dt * (-vi - mf.x);
          bi = -d.sqrt_dt * (vr - mf.y);
          const double ab = hypot(br, bi);
          if (ab > 1.0) {
            br /= ab;
            bi /= ab;
            ++ntrig;
          }
        }
        xr = z[u] - br;
        xi_ = -bi;
        // cmf: x * vbar ; cfb: xi*xbar, xbar*xbar
        cmf_r += xr * mf.x - xi_ * mf.y;
        cmf_i += xr * mf.y + xi_ * mf.x;
        s1_r += z[u] * br;
        s1_i += z[u] * bi;
        s2_r += br * br - bi * bi;
        s2_i += 2.0 * br * bi;
      }
      *reinterpret_cast<double2*>(a.XF + xf_index(d, w, n, 0)) = make_double2(xr, xi_);
      if (a.xbar_out != nullptr) a.xbar_out[(size_t)w * d.Np + n] = make_double2(br, bi);
      if (a.xs_out != nullptr) a.xs_out[(size_t)w * d.Np + n] = make_double2(xr, xi_);
    }
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    cmf_r += __shfl_xor_sync(0xffffffffu, cmf_r, m);
    cmf_i += __shfl_xor_sync(0xffffffffu, cmf_i, m);
    s1_r += __shfl_xor_sync(0xffffffffu, s1_r, m);
    s1_i += __shfl_xor_sync(0xffffffffu, s1_i, m);
    s2_r += __shfl_xor_sync(0xffffffffu, s2_r, m);
    s2_i += __shfl_xor_sync(0xffffffffu, s2_i, m);
    ntrig += __shfl_xor_sync(0xffffffffu, ntrig, m);
  }
  if (lane == 0) {
    a.cmfcfb[2 * w] = make_double2(-d.sqrt_dt * cmf_r, -d.sqrt_dt * cmf_i);
    a.cmfcfb[2 * w + 1] = make_double2(s1_r - 0.5 * s2_r, s1_i - 0.5 * s2_i);
    if (ntrig) atomicAdd(reinterpret_cast<unsigned long long*>(a.counters), (unsigned long long)ntrig);
  }
}

// ============================================================================
// active mask (qmc/afqmc.py:232)
// ============================================================================
__global__ void active_kernel(const double* __restrict__ weight, int* __restrict__ active,
                              long long* counters, Dims d) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= d.Wp) return;
  int act = 0;
  if (w < d.W) {
    act = fabs(weight[w]) > 1e-8 ? 1 : 0;
    if (!act) atomicAdd(reinterpret_cast<unsigned long long*>(counters + 2), 1ull);
  }
  active[w] = act;
}

// ============================================================================
// walkers.use_log_shift (walkers/handler.py:228,456-475): running averages of log(mean |ot|),
// log(mean |detR|) and mean |log_detR| over the whole population; determinants are reported as
// exp(logdet - log_shift) (single_det.py:159,192,320) and detR as exp(log_det - detR_shift)
// (single_det.py:250).  Every ratio the propagation uses is shift-free, so the walk itself does
// not change: the shifts only rescale the stored ot / detR / log_detR (and the Overlap column).
// ============================================================================
struct LogShifts {
  double log_shift, detR_shift, log_detR_shift;
  long long counter;  // handler.shift_counter (starts at 1)
  int enabled, pad;
};
// sums[0..2] = sum_w |ot|, sum_w |detR|, sum_w |log_detR| over this device's walkers (fixed order)
__global__ void __launch_bounds__(1024) shift_sums_kernel(const double2* __restrict__ ot, const double* __restrict__ detR,
                                                          const double* __restrict__ log_detR,
                                                          double* __restrict__ sums, int W) {
  __shared__ double red[3][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double v[3] = {0.0, 0.0, 0.0};
  for (int w = tid; w < W; w += 1024) {
    const double2 o = ot[w];
    v[0] += hypot(o.x, o.y);
    v[1] += fabs(detR[w]);
    v[2] += fabs(log_detR[w]);
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], m);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 3; ++k) {
      double x = red[k][lane];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
      if (lane == 0) sums[k] = x;
    }
  }
}
// handler.py:465-475 from the (all-reduced) sums
__global__ void shift_update_kernel(LogShifts* s, const double* __restrict__ sums, double ntot) {
  const double n = (double)s->counter, nm1 = (double)(s->counter - 1);
  s->log_shift = (s->log_shift * nm1 + log(sums[0] / ntot)) / n;
  s->detR_shift = (s->detR_shift * nm1 + log(sums[1] / ntot)) / n;
  s->log_detR_shift = (s->log_detR_shift * nm1 + sums[2] / ntot) / n;
  s->counter += 1;
}

// ============================================================================
// K6: hybrid weight update + cap (propagation/continuous.py:202-214,264-292,
//     qmc/afqmc.py:235-236)
// ============================================================================
struct WeightArgs {
  double* weight;
  double2* phase;  // free projection only
  double2* ot;
  double2* ehyb;
  const double2* ovlp_new;
  const double2* ovlp_old;  // walker.ot, or the overlap recomputed at the top of the step when ot was stale
  double2* ot_true;         // use_log_shift: the un-shifted overlap (ovlp_old of the next step)
  const LogShifts* shifts;
  double2* bpfac;           // back propagation: [W][2] running products of the phase factors I/|I| and of the
                            // cosine factors since the last reset (walkers/stack.py:51-71,118-121), or null
  const double2* cmfcfb;
  const int* active;
  const double* total_weight;
  long long* counters;
  Dims d;
  const StepParams* sp;  // eshift, step
};

__global__ void weight_kernel(WeightArgs a) {
  const Dims& d = a.d;
  const double eshift = a.sp->eshift, eshift_im = a.sp->eshift_im;
  const long long step = (long long)a.sp->step;
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= d.W) return;
  double wt = a.weight[w];
  // reported overlap: exp(-log_shift) times the determinant (single_det.py:192)
  const double osc = a.shifts->enabled ? exp(-a.shifts->log_shift) : 1.0;
  if (a.active[w] && (d.flags & FLAG_FREE_PROJECTION)) {
    // propagate_walker_free (continuous.py:194-200): the constant terms go into weight and phase
    const double2 cmf = a.cmfcfb[2 * w];
    const double er = exp(cmf.x + d.dt * eshift);
    double sn, cs;
    sincos(cmf.y + d.dt * eshift_im, &sn, &cs);
    const double magn = hypot(er * cs, er * sn);
    const double dtheta = atan2(er * sn, er * cs);
    wt = wt * magn;
    sincos(dtheta, &sn, &cs);
    const double2 ph = a.phase[w];
    a.phase[w] = make_double2(ph.x * cs - ph.y * sn, ph.x * sn + ph.y * cs);
    a.ot[w] = make_double2(a.ovlp_new[w].x * osc, a.ovlp_new[w].y * osc);
    a.ot_true[w] = a.ovlp_new[w];
  } else if (a.active[w]) {
    // ovlp_old == walker.ot: the overlap of the walker before this step (single_det.py:321 equals
    // the stored ot up to rounding; after a re-orthogonalisation ot was divided by detR)
    const double2 oo = a.ovlp_old[w], on = a.ovlp_new[w];
    const cplx ratio = cdiv({on.x, on.y}, {oo.x, oo.y});
    const double2 cmf = a.cmfcfb[2 * w], cfb = a.cmfcfb[2 * w + 1];
    // cmath.log: principal branch
    const double lr = log(hypot(ratio.re, ratio.im)), li = atan2(ratio.im, ratio.re);
    double eh_r = -(lr + cfb.x + cmf.x) / d.dt;
    const double eh_i = -(li + cfb.y + cmf.y) / d.dt;
    if (hypot(eshift, eshift_im) >= 1e-10) {
      if (eh_r > eshift + d.ebound) {
        eh_r = eshift + d.ebound;
        atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + 1), 1ull);
      } else if (eh_r < eshift - d.ebound) {
        eh_r = eshift - d.ebound;
        atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + 1), 1ull);
      }
    }
    const double2 eo = a.ehyb[w];
    // importance function exp(-dt (0.5 (Eh + Eh_old) - eshift))
    const double ar = -d.dt * (0.5 * (eh_r + eo.x) - eshift);
    const double ai = -d.dt * (0.5 * (eh_i + eo.y) - eshift_im);
    const double er = exp(ar);
    double sn, cs;
    sincos(ai, &sn, &cs);
    const double magn = hypot(er * cs, er * sn);
    a.ehyb[w] = make_double2(eh_r, eh_i);
    if (!isinf(magn)) {
      const double dtheta = -d.dt * eh_i - cfb.y;
      const double cf = fmax(0.0, cos(dtheta));
      wt = wt * (magn * cf);
      if (a.bpfac != nullptr) {
        // FieldConfig.update(xmxbar, wfac) with wfac = (I / |I|, cosine_fac), or (0, 0) when
        // |I| <= 1e-16 (continuous.py:284-289)
        const bool ok = magn > 1e-16;
        const double2 ph = a.bpfac[2 * w];
        a.bpfac[2 * w] = ok ? make_double2(ph.x * cs - ph.y * sn, ph.x * sn + ph.y * cs) : make_double2(0.0, 0.0);
        a.bpfac[2 * w + 1].x *= ok ? cf : 0.0;
      }
    } else {
      wt = 0.0;
    }
    a.ot[w] = make_double2(on.x * osc, on.y * osc);
    a.ot_true[w] = on;
  }
  if (step > 1) {
    const double cap = a.total_weight[0] * 0.10;
    if (fabs(wt) > cap) wt = cap;
  }
  a.weight[w] = wt;
}

// ============================================================================
// Multi-determinant trials (walkers/multi_det.py): the per-determinant stages are
// the single-determinant kernels run once per determinant; these kernels contract over the
// determinant index with the weights w_i = conj(c_i) <psi_i|phi>.
// ============================================================================
// ovlp[w] = sum_i conj(c_i) ovlp_det[i][w]   (multi_det.py:141-166, :198-231)
__global__ void md_overlap_kernel(const double2* __restrict__ coeffs, const double2* __restrict__ ovlp_det,
                                  size_t det_stride, int ndets, double2* __restrict__ out, int Wp) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= Wp) return;
  double re = 0.0, im = 0.0;
  for (int i = 0; i < ndets; ++i) {
    const double2 c = coeffs[i], o = ovlp_det[(size_t)i * det_stride + w];
    re += c.x * o.x + c.y * o.y;   // conj(c) * o
    im += c.x * o.y - c.y * o.x;
  }
  out[w] = make_double2(re, im);
}

// Force bias of propagation/generic.py:154-157: V[n] = sum_i w_i (X_i_up + X_i_dn)[n] / sum_i w_i,
// written as the "up" half of a combined X (the "down" half is zero) so that field_kernel is unchanged
__global__ void __launch_bounds__(256) md_x_kernel(const double2* __restrict__ coeffs,
                                                   const double2* __restrict__ ovlp_det, size_t ovlp_stride,
                                                   const double2* __restrict__ X, size_t x_stride, int ndets,
                                                   double2* __restrict__ XC, Dims d) {
  const int w = blockIdx.y;
  if (w >= d.Wp) return;
  double wr[8], wi[8];   // ndets <= PXB_MAX_DETS
  double sr = 0.0, si = 0.0;
  for (int i = 0; i < ndets; ++i) {
    const double2 c = coeffs[i], o = ovlp_det[(size_t)i * ovlp_stride + w];
    wr[i] = c.x * o.x + c.y * o.y;
    wi[i] = c.x * o.y - c.y * o.x;
    sr += wr[i];
    si += wi[i];
  }
  for (int n = blockIdx.x * blockDim.x + threadIdx.x; n < d.Np; n += gridDim.x * blockDim.x) {
    double nr = 0.0, ni = 0.0;
    for (int i = 0; i < ndets; ++i) {
      const double2* Xi = X + (size_t)i * x_stride;
      const double2 a = Xi[(size_t)w * d.Np + n], b = Xi[((size_t)d.Wp + w) * d.Np + n];
      const double vr = a.x + b.x, vi = a.y + b.y;
      nr += wr[i] * vr - wi[i] * vi;
      ni += wr[i] * vi + wi[i] * vr;
    }
    const cplx q = cdiv({nr, ni}, {sr, si});
    XC[(size_t)w * d.Np + n] = make_double2(q.re, q.im);
    XC[((size_t)d.Wp + w) * d.Np + n] = make_double2(0.0, 0.0);
  }
}

// local_energy_multi_det (estimators/mixed.py:439-448): E[w][k] = sum_i w_i E_i[w][k] / sum_i w_i.
// only_total != 0: just the total energy into out[w] (the local-energy weight update)
__global__ void md_energy_kernel(const double2* __restrict__ coeffs, const double2* __restrict__ ovlp_det,
                                 size_t ovlp_stride, const double2* __restrict__ eloc_det, size_t eloc_stride,
                                 int ndets, double2* __restrict__ out, int only_total, int W) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= W) return;
  double sr = 0.0, si = 0.0, er[3] = {0, 0, 0}, ei[3] = {0, 0, 0};
  for (int i = 0; i < ndets; ++i) {
    double wr = 1.0, wi = 0.0;
    if (ndets > 1) {
      const double2 c = coeffs[i], o = ovlp_det[(size_t)i * ovlp_stride + w];
      wr = c.x * o.x + c.y * o.y;
      wi = c.x * o.y - c.y * o.x;
    }
    sr += wr;
    si += wi;
    for (int k = 0; k < (only_total ? 1 : 3); ++k) {
      const double2 e = eloc_det[(size_t)i * eloc_stride + 3 * (size_t)w + k];
      er[k] += wr * e.x - wi * e.y;
      ei[k] += wr * e.y + wi * e.x;
    }
  }
  for (int k = 0; k < (only_total ? 1 : 3); ++k) {
    const cplx q = cdiv({er[k], ei[k]}, {sr, si});
    out[(only_total ? (size_t)w : 3 * (size_t)w + k)] = make_double2(q.re, q.im);
  }
}

// Local-energy weight update (propagation/continuous.py:216-231,294-318; row A8' of SURVEY.md):
//   eloc = walker.local_energy (Green's functions of the walker BEFORE the step, determinant
//   weights AFTER it), real part bounded to eshift +- sqrt(2/dt) once eshift != 0,
//   weight *= exp(-dt/2 Re(eloc_b + walker.eloc - eshift)) max(0, cos(arg(ot_new / ot_old)))
struct WeightLeArgs {
  double* weight;
  double2* ot;
  double2* walker_eloc;        // walker.eloc (walkers/walker.py:37), travels with the walker
  double2* ot_true;
  const LogShifts* shifts;
  const double2* eloc_mix;     // [Wp] eloc of this step
  const double2* ovlp_new;
  const double2* ovlp_old;
  const int* active;
  const double* total_weight;
  long long* counters;
  Dims d;
  const StepParams* sp;
};

__global__ void weight_le_kernel(WeightLeArgs a) {
  const Dims& d = a.d;
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w >= d.W) return;
  const double eshift = a.sp->eshift, eshift_im = a.sp->eshift_im;
  const long long step = (long long)a.sp->step;
  double wt = a.weight[w];
  if (a.active[w]) {
    const double2 oo = a.ovlp_old[w], on = a.ovlp_new[w];
    const cplx ratio = cdiv({on.x, on.y}, {oo.x, oo.y});
    const double2 el = a.eloc_mix[w], eold = a.walker_eloc[w];
    double re_b = el.x;
    if (hypot(eshift, eshift_im) >= 1e-10) {
      if (el.x > eshift + d.ebound) {
        re_b = eshift + d.ebound;
        atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + 1), 1ull);
      } else if (el.x < eshift - d.ebound) {
        re_b = eshift - d.ebound;
        atomicAdd(reinterpret_cast<unsigned long long*>(a.counters + 1), 1ull);
      }
    }
    const double magn = exp(-0.5 * d.dt * (re_b + eold.x - eshift));
    a.walker_eloc[w] = el;
    if (!isinf(magn)) {
      const double dtheta = atan2(ratio.im, ratio.re);
      wt = wt * (magn * fmax(0.0, cos(dtheta)));
    } else {
      wt = 0.0;
    }
    const double osc = a.shifts->enabled ? exp(-a.shifts->log_shift) : 1.0;
    a.ot[w] = make_double2(on.x * osc, on.y * osc);
    a.ot_true[w] = on;
  }
  if (step > 1) {
    const double cap = a.total_weight[0] * 0.10;
    if (fabs(wt) > cap) wt = cap;
  }
  a.weight[w] = wt;
}

// ============================================================================
// K7b: energy assembly (estimators/generic.py:187-189,216-221)
// ============================================================================
struct EnergyArgs {
  const double2* X;    // [2][Wp][Np]
  const double2* exx;  // [2][Wp]
  const double2* e1b;  // [Wp]
  double2* eloc;       // [W][3]
  Dims d;
};

__global__ void __launch_bounds__(256) energy_kernel(EnergyArgs a) {
  const Dims& d = a.d;
  const int lane = threadIdx.x & 31;
  const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (w >= d.W) return;
  const double2* Xa = a.X + (size_t)w * d.Np;
  const double2* Xb = a.X + ((size_t)d.Wp + w) * d.Np;
  double aa_r = 0, aa_i = 0, bb_r = 0, bb_i = 0, ab_r = 0, ab_i = 0;
  for (int n = lane; n < d.N; n += 32) {
    const double2 x = Xa[n], y = Xb[n];
    aa_r += x.x * x.x - x.y * x.y;
    aa_i += 2.0 * x.x * x.y;
    bb_r += y.x * y.x - y.y * y.y;
    bb_i += 2.0 * y.x * y.y;
    ab_r += x.x * y.x - x.y * y.y;
    ab_i += x.x * y.y + x.y * y.x;
  }
#pragma unroll
  for (int m = 16; m > 0; m >>= 1) {
    aa_r += __shfl_xor_sync(0xffffffffu, aa_r, m);
    aa_i += __shfl_xor_sync(0xffffffffu, aa_i, m);
    bb_r += __shfl_xor_sync(0xffffffffu, bb_r, m);
    bb_i += __shfl_xor_sync(0xffffffffu, bb_i, m);
    ab_r += __shfl_xor_sync(0xffffffffu, ab_r, m);
    ab_i += __shfl_xor_sync(0xffffffffu, ab_i, m);
  }
  if (lane == 0) {
    const double ec_r = aa_r + bb_r + 2.0 * ab_r, ec_i = aa_i + bb_i + 2.0 * ab_i;
    const double2 xa = a.exx[w], xb = a.exx[(size_t)d.Wp + w], e1 = a.e1b[w];
    const double e2_r = 0.5 * (ec_r - (xa.x + xb.x)), e2_i = 0.5 * (ec_i - (xa.y + xb.y));
    a.eloc[3 * (size_t)w + 0] = make_double2(e1.x + e2_r + d.ecore, e1.y + e2_i);
    a.eloc[3 * (size_t)w + 1] = make_double2(e1.x + d.ecore, e1.y);
    a.eloc[3 * (size_t)w + 2] = make_double2(e2_r, e2_i);
  }
}

// ============================================================================
// K9: Mixed.update accumulation (estimators/mixed.py:211-225).  One CTA, fixed
// reduction tree (deterministic).  estimates: complex[10], enum order
// uweight, weight, enumer, edenom, eproj, e1b, e2b, ehyb, ovlp, time.
// ============================================================================
struct AccArgs {
  const double* weight;
  const double2* phase;
  const double* unscaled;
  const double2* ot;
  const double2* ehyb;
  const double2* eloc;
  double2* estimates;
  Dims d;
  int with_energy;
};

__global__ void __launch_bounds__(1024) accumulate_kernel(AccArgs a) {
  __shared__ double red[9][32];
  const Dims& d = a.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double v[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int w = tid; w < d.W; w += 1024) {
    const double wt = a.weight[w];
    if (a.with_energy) {
      v[0] += wt * a.eloc[3 * (size_t)w].x;      // enumer
      v[1] += wt * a.eloc[3 * (size_t)w + 1].x;  // e1b
      v[2] += wt * a.eloc[3 * (size_t)w + 2].x;  // e2b
      v[3] += wt;                                // edenom
    }
    v[4] += a.unscaled[w];
    v[5] += wt;
    const double2 o = a.ot[w], e = a.ehyb[w];
    v[6] += wt * hypot(o.x, o.y);
    v[7] += wt * e.x;
    v[8] += wt * e.y;
  }
#pragma unroll
  for (int k = 0; k < 9; ++k) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], m);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 9; ++k) {
      double x = red[k][lane];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
      v[k] = x;
    }
    if (lane == 0) {
      a.estimates[2].x += v[0];
      a.estimates[5].x += v[1];
      a.estimates[6].x += v[2];
      a.estimates[3].x += v[3];
      a.estimates[0].x += v[4];
      a.estimates[1].x += v[5];
      a.estimates[8].x += v[6];
      a.estimates[7].x += v[7];
      a.estimates[7].y += v[8];
    }
  }
}

// Mixed one-body density matrix (estimators/mixed.py:226-229): sum_w w Re(G_w) with
// G_w = conj(psi) Theta_w is conj(psi) (sum_w w Theta_w), so the device only reduces Theta over the
// walkers: out[i][p] += sum_w weight_w Theta_w[i][p].  One CTA per (orbital i, basis chunk kc) row
// of the OF layout: every walker group contributes one contiguous 256-byte fragment.  Fixed
// reduction order (deterministic).
__global__ void __launch_bounds__(128) theta_wsum_kernel(const double* __restrict__ theta,
                                                         const double* __restrict__ weight,
                                                         double2* __restrict__ out, Dims d) {
  __shared__ double2 red[128];
  const int row = blockIdx.x;  // i * KC + kc
  const int i = row / d.KC, kc = row % d.KC;
  const int e = threadIdx.x & 15, sub = threadIdx.x >> 4;  // element (wl, t) of a fragment, 8 groups at a time
  const int wl = e >> 2;
  double re = 0.0, im = 0.0;
  for (int wg = sub; wg < d.WG; wg += 8) {
    const int w = 4 * wg + wl;
    if (w < d.W) {
      const double2 v = *reinterpret_cast<const double2*>(theta + (((size_t)wg * d.ne + i) * d.KC + kc) * 32 + e * 2);
      const double wt = weight[w];
      re += wt * v.x;
      im += wt * v.y;
    }
  }
  red[threadIdx.x] = make_double2(re, im);
  __syncthreads();
  if (threadIdx.x < 4) {  // thread t sums its 4 walkers-in-group x 8 sub-slices in a fixed order
    double sr = 0.0, si = 0.0;
    for (int s2 = 0; s2 < 8; ++s2)
      for (int l = 0; l < 4; ++l) {
        const double2 v = red[s2 * 16 + l * 4 + threadIdx.x];
        sr += v.x;
        si += v.y;
      }
    const int p = 4 * kc + threadIdx.x;
    if (p < d.M) {
      out[(size_t)i * d.M + p].x += sr;
      out[(size_t)i * d.M + p].y += si;
    }
  }
}

// free projection (estimators/mixed.py:151-177): every sum carries the complex factor
// wfac = weight * ot * phase
__global__ void __launch_bounds__(1024) accumulate_free_kernel(AccArgs a) {
  __shared__ double red[14][32];
  const Dims& d = a.d;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  double v[14];
#pragma unroll
  for (int k = 0; k < 14; ++k) v[k] = 0.0;
  for (int w = tid; w < d.W; w += 1024) {
    const double wt = a.weight[w];
    const double2 o = a.ot[w], ph = a.phase[w];
    const double fr = wt * o.x, fi = wt * o.y;          // weight * ot
    const double wr = fr * ph.x - fi * ph.y, wi = fr * ph.y + fi * ph.x;  // ... * phase
    if (a.with_energy) {
#pragma unroll
      for (int k = 0; k < 3; ++k) {
        const double2 e = a.eloc[3 * (size_t)w + k];
        v[2 * k] += wr * e.x - wi * e.y;
        v[2 * k + 1] += wr * e.y + wi * e.x;
      }
      v[6] += wr;
      v[7] += wi;
    }
    v[8] += a.unscaled[w];
    v[9] += wr;
    v[10] += wi;
    const double2 eh = a.ehyb[w];
    v[11] += wr * eh.x - wi * eh.y;
    v[12] += wr * eh.y + wi * eh.x;
    v[13] += wt * hypot(o.x, o.y);
  }
#pragma unroll
  for (int k = 0; k < 14; ++k) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], m);
    if (lane == 0) red[k][warp] = v[k];
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 14; ++k) {
      double x = red[k][lane];
#pragma unroll
      for (int m = 16; m > 0; m >>= 1) x += __shfl_xor_sync(0xffffffffu, x, m);
      v[k] = x;
    }
    if (lane == 0) {
      a.estimates[2].x += v[0];
      a.estimates[2].y += v[1];
      a.estimates[5].x += v[2];
      a.estimates[5].y += v[3];
      a.estimates[6].x += v[4];
      a.estimates[6].y += v[5];
      a.estimates[3].x += v[6];
      a.estimates[3].y += v[7];
      a.estimates[0].x += v[8];
      a.estimates[1].x += v[9];
      a.estimates[1].y += v[10];
      a.estimates[7].x += v[11];
      a.estimates[7].y += v[12];
      a.estimates[8].x += v[13];
    }
  }
}

// ============================================================================
// K10: population control (walkers/handler.py:225-338)
// ============================================================================
__global__ void abs_weight_kernel(const double* __restrict__ weight, double* __restrict__ out, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) out[i] = fabs(weight[i]);
}

// sequential (index-order) running sum of m shared-memory values by ONE thread; the loads of a
// block of 16 are issued before the dependent DADD chain so that only the adds are serial
template <bool STORE>
__device__ __forceinline__ double serial_sum16(double run, double* chunk, int m) {
  int i = 0;
  for (; i + 16 <= m; i += 16) {
    double v[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) v[j] = chunk[i + j];
#pragma unroll
    for (int j = 0; j < 16; ++j) {
      run = __dadd_rn(run, v[j]);
      v[j] = run;
    }
    if (STORE) {
#pragma unroll
      for (int j = 0; j < 16; ++j) chunk[i + j] = v[j];
    }
  }
  for (; i < m; ++i) {
    run = __dadd_rn(run, chunk[i]);
    if (STORE) chunk[i] = run;
  }
  return run;
}

// total = python-style sequential sum of the global |weights| (handler.py:233);
// gws = gw / scale (handler.py:251); local weights rescaled (handler.py:247-249).
// One CTA; the sum itself is done by thread 0 in index order.
struct RescaleArgs {
  const double* gw;     // [Wtot]
  double* gws;          // [Wtot] scratch: rescaled global weights
  double* weight;       // local [W]
  double* unscaled;     // local [W]
  double* total_weight; // [1]
  long long* counters;
  int W, Wtot;
  int write_local;      // 0: leave weight / unscaled alone (peer-memory comb: peers may still read them)
};

__global__ void __launch_bounds__(1024) pop_rescale_kernel(RescaleArgs a) {
  __shared__ double chunk[1024];
  __shared__ double s_total;
  const int tid = threadIdx.x;
  double total = 0.0;  // only meaningful in thread 0
  for (int base = 0; base < a.Wtot; base += 1024) {
    if (base + tid < a.Wtot) chunk[tid] = a.gw[base + tid];
    __syncthreads();
    if (tid == 0) total = serial_sum16<false>(total, chunk, min(1024, a.Wtot - base));
    __syncthreads();
  }
  if (tid == 0) s_total = total;
  __syncthreads();
  total = s_total;
  if (total < 1e-8 || a.counters[4] != 0) {
    // the reference exits here (handler.py:236-241); counters[4] is a STICKY flag that nothing else
    // writes: the comb plan, the copies and the weight reset all become no-ops once it is set, and
    // the driver raises when it polls it (Walkers.check_total_weight)
    if (tid == 0) a.counters[4] = 1;
    for (int i = tid; i < a.Wtot; i += 1024) a.gws[i] = 0.0;
    return;
  }
  const double scale = __ddiv_rn(total, (double)a.Wtot);
  if (tid == 0) a.total_weight[0] = total;
  for (int i = tid; i < a.Wtot; i += 1024) a.gws[i] = __ddiv_rn(a.gw[i], scale);
  if (!a.write_local) return;
  for (int i = tid; i < a.W; i += 1024) {
    const double wt = a.weight[i];
    a.unscaled[i] = wt;
    a.weight[i] = __ddiv_rn(wt, scale);
  }
}

// comb selection (handler.py:271-301).  cprobs = sequential cumsum; every tooth
// finds the first walker with tooth < cprobs[iw] (identical to the reference's
// two-pointer sweep because both sequences are non-decreasing).
struct CombArgs {
  const double* gws;  // [Wtot] rescaled weights
  double* cprobs;     // [Wtot] scratch
  int* parent_ix;     // [Wtot]
  int* pairs;         // [1 + 2*Wtot]
  long long* counters;
  int Wtot;
  const StepParams* sp;  // comb_r
};

// number of comb teeth (i + r) * spacing, i in [0, n), that lie below c -- evaluated with the
// reference's own floating-point expression around the analytic estimate, so the result is
// exactly what the two-pointer sweep of handler.py:277-286 counts
__device__ __forceinline__ int teeth_below(double c, double r, double spacing, int n) {
  double e = floor(c / spacing - r);
  int ic = e < 0.0 ? 0 : (e > (double)n ? n : (int)e);
  while (ic > 0 && !(__dmul_rn(__dadd_rn((double)(ic - 1), r), spacing) < c)) --ic;
  while (ic < n && __dmul_rn(__dadd_rn((double)ic, r), spacing) < c) ++ic;
  return ic;
}

__global__ void __launch_bounds__(1024) comb_plan_kernel(CombArgs a) {
  __shared__ double chunk[1024];
  __shared__ double s_total;
  __shared__ int s_warp[2][32];
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, n = a.Wtot;
  const double r = a.sp->comb_r;
  double run = 0.0;
  for (int base = 0; base < n; base += 1024) {
    if (base + tid < n) chunk[tid] = a.gws[base + tid];
    __syncthreads();
    if (tid == 0) run = serial_sum16<true>(run, chunk, min(1024, n - base));
    __syncthreads();
    if (base + tid < n) a.cprobs[base + tid] = chunk[tid];
    __syncthreads();
  }
  if (tid == 0) s_total = run;
  __syncthreads();
  __threadfence_block();
  const double spacing = __ddiv_rn(s_total, (double)n);
  if (!(spacing > 0.0) || a.counters[4] != 0) {  // vanished population (sticky flag): nothing to plan
    for (int i = tid; i < n; i += 1024) a.parent_ix[i] = 1;
    if (tid == 0) a.pairs[0] = 0;
    return;
  }
  // parent_ix[iw] = teeth in [cprobs[iw-1], cprobs[iw]); teeth past the last walker stay with it
  // (the reference would raise IndexError there).  Each thread owns a contiguous segment, so the
  // kill (parent == 0) / clone (parent > 1) lists come out ascending, zipped position-wise.
  const int seg = (n + 1023) / 1024;
  const int i0 = min(tid * seg, n), i1 = min(i0 + seg, n);
  int nk = 0, nc = 0;
  {
    int below = (i0 == 0) ? 0 : teeth_below(a.cprobs[i0 - 1], r, spacing, n);
    for (int i = i0; i < i1; ++i) {
      const int upto = (i == n - 1) ? n : teeth_below(a.cprobs[i], r, spacing, n);
      const int p = upto - below;
      below = upto;
      a.parent_ix[i] = p;
      nk += (p == 0);
      nc += (p > 1);
    }
  }
  // exclusive block scan of (nk, nc)
  int sk = nk, sc = nc;
#pragma unroll
  for (int off = 1; off < 32; off <<= 1) {
    const int vk = __shfl_up_sync(0xffffffffu, sk, off), vc = __shfl_up_sync(0xffffffffu, sc, off);
    if (lane >= off) {
      sk += vk;
      sc += vc;
    }
  }
  if (lane == 31) {
    s_warp[0][warp] = sk;
    s_warp[1][warp] = sc;
  }
  __syncthreads();
  if (warp == 0) {
    int wk = s_warp[0][lane], wc = s_warp[1][lane];
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const int vk = __shfl_up_sync(0xffffffffu, wk, off), vc = __shfl_up_sync(0xffffffffu, wc, off);
      if (lane >= off) {
        wk += vk;
        wc += vc;
      }
    }
    s_warp[0][lane] = wk;
    s_warp[1][lane] = wc;
  }
  __syncthreads();
  int ok = sk - nk + (warp ? s_warp[0][warp - 1] : 0);
  int oc = sc - nc + (warp ? s_warp[1][warp - 1] : 0);
  for (int i = i0; i < i1; ++i) {
    const int p = a.parent_ix[i];
    if (p == 0) a.pairs[1 + 2 * (ok++) + 1] = i;
    if (p > 1) a.pairs[1 + 2 * (oc++)] = i;
  }
  if (tid == 0) {
    const int np = min(s_warp[0][31], s_warp[1][31]);
    a.pairs[0] = np;
    a.counters[3] += np;
  }
}

// payload copy: phi (OF layout) + per-walker scalars
struct CopyArgs {
  double* phi;
  double* theta;   // rotated Green's function travels with the walker (stays valid)
  double2* e1b;
  double* weight;
  double* unscaled;
  double2* ot;
  double2* ehyb;
  double2* eloc;
  double* detR;
  double* log_detR;
  double2* phase;
  double2* weloc;   // walker.eloc of the local-energy weight update
  double2* ottrue;  // un-shifted overlap (use_log_shift)
  double2* bpfac;   // [W][2] back-propagation weight-factor products
  double2* X;       // [2][Wp][Np] X_s = R_s^T Theta_s travels with the walker like Theta does
  double* phi_old;  // back propagation only (else nullptr): walker.phi_old, OF layout
  double* fc;       // back propagation only: field history, rows = nbp * NKC per walker group
  int fc_rows;
  Dims d;
};

__device__ __host__ __forceinline__ size_t payload_doubles(const Dims& d, bool bp, int fc_rows) {
  return (size_t)(bp ? 3 : 2) * d.ne * d.KC * 8 + (bp ? (size_t)fc_rows * 8 : 0) + (size_t)4 * d.Np + 26;
}

// X rows of walker src (array sX) -> walker dst (array dX)
__device__ __forceinline__ void copy_x(double2* dX, const double2* sX, const Dims& d, int dst, int src) {
  for (int idx = threadIdx.x; idx < 2 * d.Np; idx += blockDim.x) {
    const int s = idx / d.Np, n = idx - s * d.Np;
    dX[((size_t)s * d.Wp + dst) * d.Np + n] = sX[((size_t)s * d.Wp + src) * d.Np + n];
  }
}

// walker-interleaved arrays [W/4][rows][4 walkers][8 doubles]: copy the `rows` 64-byte pieces of
// walker src (in sbase) to walker dst (in dbase)
__device__ __forceinline__ void copy_rows(double* dbase, const double* sbase, int rows, int dst, int src) {
  for (int idx = threadIdx.x; idx < rows * 4; idx += blockDim.x) {
    const int r = idx >> 2, q = idx & 3;
    const size_t so = ((size_t)(src >> 2) * rows + r) * 32 + (src & 3) * 8 + q * 2;
    const size_t dn = ((size_t)(dst >> 2) * rows + r) * 32 + (dst & 3) * 8 + q * 2;
    *reinterpret_cast<double2*>(dbase + dn) = *reinterpret_cast<const double2*>(sbase + so);
  }
}
// same between walker w of an interleaved array and a contiguous buffer of rows * 8 doubles
__device__ __forceinline__ void pack_rows(double* base, double* buf, int rows, int w, int unpack) {
  for (int idx = threadIdx.x; idx < rows * 4; idx += blockDim.x) {
    const int r = idx >> 2, q = idx & 3;
    double2* g = reinterpret_cast<double2*>(base + ((size_t)(w >> 2) * rows + r) * 32 + (w & 3) * 8 + q * 2);
    double2* l = reinterpret_cast<double2*>(buf + (size_t)r * 8 + q * 2);
    if (unpack)
      *g = *l;
    else
      *l = *g;
  }
}

// pairs: device list [1 + 2*n] of GLOBAL indices; only pairs with both ends on
// this device (offset <= idx < offset + W) are copied here
__global__ void __launch_bounds__(256) copy_pairs_kernel(CopyArgs a, const int* pairs, int offset) {
  const Dims& d = a.d;
  const int np = pairs[0];
  for (int pi = blockIdx.x; pi < np; pi += gridDim.x) {
    const int src = pairs[1 + 2 * pi] - offset, dst = pairs[2 + 2 * pi] - offset;
    if (src < 0 || src >= d.W || dst < 0 || dst >= d.W) continue;
    const int n8 = d.ne * d.KC;
    copy_rows(a.phi, a.phi, n8, dst, src);
    copy_rows(a.theta, a.theta, n8, dst, src);
    copy_x(a.X, a.X, d, dst, src);
    if (a.phi_old != nullptr) {
      copy_rows(a.phi_old, a.phi_old, n8, dst, src);
      copy_rows(a.fc, a.fc, a.fc_rows, dst, src);
    }
    if (threadIdx.x == 0) {
      a.e1b[dst] = a.e1b[src];
      a.weight[dst] = a.weight[src];
      a.unscaled[dst] = a.unscaled[src];
      a.ot[dst] = a.ot[src];
      a.ehyb[dst] = a.ehyb[src];
      a.phase[dst] = a.phase[src];
      a.weloc[dst] = a.weloc[src];
      a.ottrue[dst] = a.ottrue[src];
      a.bpfac[2 * dst] = a.bpfac[2 * src];
      a.bpfac[2 * dst + 1] = a.bpfac[2 * src + 1];
      a.detR[dst] = a.detR[src];
      a.log_detR[dst] = a.log_detR[src];
      for (int k = 0; k < 3; ++k) a.eloc[3 * (size_t)dst + k] = a.eloc[3 * (size_t)src + k];
    }
  }
}

__global__ void __launch_bounds__(256) copy_list_kernel(CopyArgs a, const int* src_l, const int* dst_l,
                                                        int n) {
  const Dims& d = a.d;
  for (int pi = blockIdx.x; pi < n; pi += gridDim.x) {
    const int src = src_l[pi], dst = dst_l[pi];
    const int n8 = d.ne * d.KC;
    copy_rows(a.phi, a.phi, n8, dst, src);
    copy_rows(a.theta, a.theta, n8, dst, src);
    copy_x(a.X, a.X, d, dst, src);
    if (a.phi_old != nullptr) {
      copy_rows(a.phi_old, a.phi_old, n8, dst, src);
      copy_rows(a.fc, a.fc, a.fc_rows, dst, src);
    }
    if (threadIdx.x == 0) {
      a.e1b[dst] = a.e1b[src];
      a.weight[dst] = a.weight[src];
      a.unscaled[dst] = a.unscaled[src];
      a.ot[dst] = a.ot[src];
      a.ehyb[dst] = a.ehyb[src];
      a.phase[dst] = a.phase[src];
      a.weloc[dst] = a.weloc[src];
      a.ottrue[dst] = a.ottrue[src];
      a.bpfac[2 * dst] = a.bpfac[2 * src];
      a.bpfac[2 * dst + 1] = a.bpfac[2 * src + 1];
      a.detR[dst] = a.detR[src];
      a.log_detR[dst] = a.log_detR[src];
      for (int k = 0; k < 3; ++k) a.eloc[3 * (size_t)dst + k] = a.eloc[3 * (size_t)src + k];
    }
  }
}

// Comb data movement over peer memory (handler.py:301-334 without the Isend/Recv): every
// device runs the same plan; for each (clone, kill) pair whose KILL slot lives here the clone's
// payload is read straight out of the owning device's arena (NVLink P2P loads through the
// IPC-mapped base pointers; plain loads when the clone is local) and written into the local
// slot.  Clone slots (parent_ix > 1) are never kill slots (parent_ix == 0), so sources are
// read-only during this phase on every device; the caller brackets the phase with collectives
// (all-gather of the weights before, a barrier after).  weight[dst] receives the clone's RAW
// weight: pop_finish_kernel turns it into unscaled_weight afterwards.
constexpr int PXB_MAX_PEERS = 16;
struct PeerArgs {
  const unsigned char* base[PXB_MAX_PEERS];  // arena base of every rank (own entry: local arena)
  int rank, nranks, nw;
};

template <class T>
__device__ __forceinline__ const T* rebase(const T* local, const unsigned char* local_base,
                                           const unsigned char* peer_base) {
  return reinterpret_cast<const T*>(peer_base + (reinterpret_cast<const unsigned char*>(local) - local_base));
}

__global__ void __launch_bounds__(256) pull_pairs_kernel(CopyArgs a, PeerArgs p, const int* pairs) {
  const Dims& d = a.d;
  const int np = pairs[0];
  const unsigned char* lb = p.base[p.rank];
  for (int pi = blockIdx.x; pi < np; pi += gridDim.x) {
    const int c = pairs[1 + 2 * pi], k = pairs[2 + 2 * pi];
    if (k / p.nw != p.rank) continue;
    const int dst = k - p.rank * p.nw;
    const int sr = c / p.nw, src = c - sr * p.nw;
    const unsigned char* pb = p.base[sr];
    const double* sphi = rebase(a.phi, lb, pb);
    const double* sth = rebase(a.theta, lb, pb);
    const int n8 = d.ne * d.KC;
    copy_rows(a.phi, sphi, n8, dst, src);
    copy_rows(a.theta, sth, n8, dst, src);
    copy_x(a.X, rebase(a.X, lb, pb), d, dst, src);
    if (a.phi_old != nullptr) {
      copy_rows(a.phi_old, rebase(a.phi_old, lb, pb), n8, dst, src);
      copy_rows(a.fc, rebase(a.fc, lb, pb), a.fc_rows, dst, src);
    }
    if (threadIdx.x == 0) {
      a.e1b[dst] = rebase(a.e1b, lb, pb)[src];
      a.weight[dst] = rebase(a.weight, lb, pb)[src];
      a.ot[dst] = rebase(a.ot, lb, pb)[src];
      a.ehyb[dst] = rebase(a.ehyb, lb, pb)[src];
      a.phase[dst] = rebase(a.phase, lb, pb)[src];
      a.weloc[dst] = rebase(a.weloc, lb, pb)[src];
      a.ottrue[dst] = rebase(a.ottrue, lb, pb)[src];
      a.bpfac[2 * dst] = rebase(a.bpfac, lb, pb)[2 * src];
      a.bpfac[2 * dst + 1] = rebase(a.bpfac, lb, pb)[2 * src + 1];
      a.detR[dst] = rebase(a.detR, lb, pb)[src];
      a.log_detR[dst] = rebase(a.log_detR, lb, pb)[src];
      const double2* se = rebase(a.eloc, lb, pb);
      for (int j = 0; j < 3; ++j) a.eloc[3 * (size_t)dst + j] = se[3 * (size_t)src + j];
    }
  }
}

// after the barrier that follows pull_pairs_kernel: unscaled_weight = weight (handler.py:247-248,
// copied clone -> kill by the comb), then every weight = value (handler.py:337-338)
__global__ void pop_finish_kernel(double* weight, double* unscaled, double value, int n,
                                  const long long* counters) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (counters[4] != 0) return;  // vanished population: the weights stay as they are
  if (i < n) {
    unscaled[i] = weight[i];
    weight[i] = value;
  }
}

// pack / unpack for transfers between devices: buffer[n][payload_doubles]
__global__ void __launch_bounds__(256) pack_kernel(CopyArgs a, const int* slots, int n, double* buf,
                                                   int unpack) {
  const Dims& d = a.d;
  const bool bp = a.phi_old != nullptr;
  const size_t pd = payload_doubles(d, bp, a.fc_rows);
  for (int pi = blockIdx.x; pi < n; pi += gridDim.x) {
    const int w = slots[pi];
    double* b = buf + (size_t)pi * pd;
    const int n8 = d.ne * d.KC;
    pack_rows(a.phi, b, n8, w, unpack);
    pack_rows(a.theta, b + (size_t)n8 * 8, n8, w, unpack);
    if (bp) {
      pack_rows(a.phi_old, b + (size_t)2 * n8 * 8, n8, w, unpack);
      pack_rows(a.fc, b + (size_t)3 * n8 * 8, a.fc_rows, w, unpack);
    }
    {
      double2* xb = reinterpret_cast<double2*>(b + pd - 26 - (size_t)4 * d.Np);
      for (int idx = threadIdx.x; idx < 2 * d.Np; idx += blockDim.x) {
        const int sp = idx / d.Np, n = idx - sp * d.Np;
        double2* gx = a.X + ((size_t)sp * d.Wp + w) * d.Np + n;
        if (unpack)
          *gx = xb[idx];
        else
          xb[idx] = *gx;
      }
    }
    if (threadIdx.x == 0) {
      double* s = b + pd - 26;
      if (unpack) {
        a.weight[w] = s[0];
        a.unscaled[w] = s[1];
        a.ot[w] = make_double2(s[2], s[3]);
        a.ehyb[w] = make_double2(s[4], s[5]);
        a.detR[w] = s[6];
        a.log_detR[w] = s[7];
        for (int k = 0; k < 3; ++k) a.eloc[3 * (size_t)w + k] = make_double2(s[8 + 2 * k], s[9 + 2 * k]);
        a.e1b[w] = make_double2(s[14], s[15]);
        a.phase[w] = make_double2(s[16], s[17]);
        a.weloc[w] = make_double2(s[18], s[19]);
        a.ottrue[w] = make_double2(s[20], s[21]);
        a.bpfac[2 * w] = make_double2(s[22], s[23]);
        a.bpfac[2 * w + 1] = make_double2(s[24], s[25]);
      } else {
        s[0] = a.weight[w];
        s[1] = a.unscaled[w];
        s[2] = a.ot[w].x;
        s[3] = a.ot[w].y;
        s[4] = a.ehyb[w].x;
        s[5] = a.ehyb[w].y;
        s[6] = a.detR[w];
        s[7] = a.log_detR[w];
        for (int k = 0; k < 3; ++k) {
          s[8 + 2 * k] = a.eloc[3 * (size_t)w + k].x;
          s[9 + 2 * k] = a.eloc[3 * (size_t)w + k].y;
        }
        s[14] = a.e1b[w].x;
        s[15] = a.e1b[w].y;
        s[16] = a.phase[w].x;
        s[17] = a.phase[w].y;
        s[18] = a.weloc[w].x;
        s[19] = a.weloc[w].y;
        s[20] = a.ottrue[w].x;
        s[21] = a.ottrue[w].y;
        s[22] = a.bpfac[2 * w].x;
        s[23] = a.bpfac[2 * w].y;
        s[24] = a.bpfac[2 * w + 1].x;
        s[25] = a.bpfac[2 * w + 1].y;
      }
    }
  }
}

// skip_flag (optional): nothing is written when *skip_flag != 0 (vanished population)
// bpfac[w] = (1, 1): FieldConfig.reset (walkers/stack.py:122-127)
__global__ void bpfac_reset_kernel(double2* bpfac, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    bpfac[2 * i] = make_double2(1.0, 0.0);
    bpfac[2 * i + 1] = make_double2(1.0, 0.0);
  }
}

__global__ void fill_kernel(double* p, double v, int n, const long long* skip_flag) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (skip_flag != nullptr && *skip_flag != 0) return;
  if (i < n) p[i] = v;
}

__global__ void init_scalars_kernel(double* weight, double* unscaled, double2* ot, const double2* ovlp,
                                    double2* ehyb, double* detR, double* log_detR, double* total_weight,
                                    double2* phase, double2* weloc, double2* ottrue, double total, Dims d) {
  const int w = blockIdx.x * blockDim.x + threadIdx.x;
  if (w == 0) total_weight[0] = total;
  if (w >= d.Wp) return;
  const bool real = w < d.W;
  weight[w] = real ? 1.0 : 0.0;
  unscaled[w] = real ? 1.0 : 0.0;
  ot[w] = ovlp[w];
  ottrue[w] = ovlp[w];
  ehyb[w] = make_double2(0.0, 0.0);
  phase[w] = make_double2(1.0, 0.0);
  weloc[w] = make_double2(0.0, 0.0);
  detR[w] = 1.0;
  log_detR[w] = 0.0;
}

}  // namespace pxb

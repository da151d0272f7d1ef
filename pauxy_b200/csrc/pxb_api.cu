// C-ABI of the B200 AFQMC hot path (see include/pauxy_b200.h).
#include "../../include/pauxy_b200.h"

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <vector>

#include "pxb_common.cuh"
#include "pxb_bp.cuh"
#include "pxb_eri.cuh"
#include "pxb_exchange.cuh"
#include "pxb_gemm.cuh"
#include "pxb_greens.cuh"
#include "pxb_greens2.cuh"
#include "pxb_qr.cuh"
#include "pxb_small.cuh"
#include "pxb_taylor.cuh"
#include "pxb_taylor2.cuh"
#include "pxb_taylor3.cuh"

using namespace pxb;

namespace {

struct Region {
  size_t off = 0, bytes = 0;
  size_t stride = 0;  // per-determinant regions: bytes between the copies of consecutive determinants
};

enum ArenaId {
  A_LF, A_RF, A_BF, A_PSIT, A_H1ROT, A_VBAR,
  A_PHI_A, A_PHI_B, A_THETA, A_X, A_XF, A_VF,
  A_EXX, A_KF0, A_KF1, A_EPART, A_RTMAP, A_OB, A_E1B, A_OVLP_OLD, A_ACTIVE, A_GW, A_GWS, A_CPROBS, A_FLAG, A_PF, A_SLOG, A_E1BP, A_QRLD, A_QRMASK,
  A_FC, A_PHI_OLD, A_PHI_BP, A_PHI_BP2, A_THETA_BP, A_BP_PART, A_PSI_NAT, A_INIT_NAT, A_BFT, A_STEP_PARAMS,
  A_OVLP_DET, A_ELOC_DET, A_XC, A_COEFF, A_ELOC_MIX, A_BF2, A_PHI_STACK, A_OT_TRUE, A_BPFAC, A_BPW,
  A_THETA_STACK, A_XF2, A_KF0I, A_KF1I, A_EPART2,
  A_FIELD0,  // public fields follow: A_FIELD0 + pxb_field_id
  A_COUNT = A_FIELD0 + PXB_F_COUNT
};

}  // namespace

struct pxb_context {
  pxb_config cfg;
  Dims d;
  Region reg[A_COUNT];
  size_t arena_bytes = 0;
  unsigned char* arena = nullptr;
  bool ham_set = false;
  int phi_cur = 0;  // which of PHI_A / PHI_B holds the walkers
  int sm_count = 148;
  int reserved_sms = 0;
  int max_smem_optin = 0;
  long long launches = 0;  // kernels launched through this handle
  // Theta / overlap / e1b (A_THETA, A_E1B) correspond to the current walkers; X to the current Theta
  bool theta_valid = false, x_valid = false;
  bool eloc_valid = false;  // ELOC holds the local energies of the current walkers (travels with them)
  // optional per-stage timing with CUDA events on the launch stream (pxb_profile / pxb_stage_times)
  bool prof = false;
  struct Pending { int stage; cudaEvent_t a, b; };
  std::vector<Pending> pending;
  std::vector<cudaEvent_t> evpool;
  double stage_ms[PXB_STAGE_COUNT] = {0};
  long long stage_calls[PXB_STAGE_COUNT] = {0};
  bool greens_split = true;  // batched overlap GEMM + warp-per-walker inverse/Theta (PXB_GREENS=fused: one CTA per walker)
  bool vhs_sym = false;  // L symmetric in (p,q): the VHS GEMM computes the upper triangle only
  bool vhs_sym_allowed = true;
  bool qr_cholesky = true;  // CholeskyQR2 where the shape allows it (PXB_QR=mgs in experiment builds: Gram-Schmidt only)
  bool hs_near_sym = false;  // L symmetric in (p,q) to rounding (needed by the back propagation)
  int rtu = 0;           // row tiles kept in that case
  bool taylor_tma = true;  // persistent TMA-fed Taylor kernel (PXB_TAYLOR=direct selects the per-walker-CTA one)
  bool taylor_3m = true;   // 3-product planar kernel where the shape allows (PXB_TAYLOR=4m: taylor2_kernel)
  bool exx_eri = false;  // exchange through the half-rotated ERI quadratic form (pxb_eri.cuh)
  bool kf_shared[PXB_MAX_DETS] = {false};  // both spins use the K of spin 0 (identical half-rotated Cholesky blocks)
  bool ham_base_set = false;
  int eri_nslot = 0;
  // back propagation: stored field configurations per walker (walkers/stack.py:5-127)
  int nbp = 0;      // capacity (steps), 0: back propagation off
  int bp_step = 0;  // configurations stored since the last pxb_bp_reset
  int bp_chunks = 0;
  int bp_restore = 0;  // estimator weights: 0 BP-PhL, 1 restore_weights = 'partial', 2 'full'
  // peer-memory population control: arena bases of all ranks mapped through CUDA IPC
  int peer_rank = -1, peer_n = 0;
  unsigned char* peer_base[PXB_MAX_PEERS] = {nullptr};
  void* peer_map[PXB_MAX_PEERS] = {nullptr};  // what cudaIpcOpenMemHandle returned (to close)
  // fused driver step (pxb_step): the launch sequence of one step, captured as a CUDA graph the
  // second time a variant is seen and replayed afterwards; the scalars of the step are written to
  // A_STEP_PARAMS by one setter launch in front of it
  double eshift_im = 0.0;  // pxb_set_eshift_imag
  bool log_shift_on = false;  // walkers.use_log_shift
  bool in_step = false;  // scalars already set by pxb_step: the entry points it calls leave them alone
  bool graphs_enabled = true;
  cudaStream_t side = nullptr;  // the comb plan runs here beside the local energy
  cudaStream_t main = nullptr;  // carries the graph when the caller's stream is a default stream
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_in = nullptr, ev_out = nullptr;
  struct StepGraph {
    unsigned long long key;
    const double* xi;
    int seen;               // direct runs so far (the first one warms module loading / attributes)
    cudaGraphExec_t exec;   // nullptr until captured
    long long kernels;      // kernel launches one replay stands for
    bool theta_valid, x_valid, eloc_valid;  // state after the sequence
  };
  std::vector<StepGraph> graphs;
  long long graph_replays = 0;
  std::string err;

  // multi-determinant trial (walkers/multi_det.py): the per-determinant operands and intermediates
  // (psi, half-rotated Cholesky / ERI / one-body integrals, Theta, X, exchange, e1b, overlaps, local
  // energies) exist once per determinant; `det` selects the copy the stage launchers work on
  int ndets = 1, det = 0;
  unsigned dets_set = 0;  // bit i: pxb_set_hamiltonian / pxb_set_trial_det has supplied determinant i
  template <class T>
  T* ptr(int id) const {
    return reinterpret_cast<T*>(arena + reg[id].off + (size_t)det * reg[id].stride);
  }
  template <class T>
  T* ptr0(int id) const {  // determinant 0 / base of a per-determinant region
    return reinterpret_cast<T*>(arena + reg[id].off);
  }
  template <class T>
  T* field(int fid) const {
    return ptr<T>(A_FIELD0 + fid);
  }
  double* phi() const { return ptr<double>(phi_cur ? A_PHI_B : A_PHI_A); }
  double* phi_other() const { return ptr<double>(phi_cur ? A_PHI_A : A_PHI_B); }
};

namespace {

int fail(pxb_handle h, int code, const std::string& msg) {
  if (h) h->err = msg;
  return code;
}

#define PXB_CUDA(h, expr)                                                              \
  do {                                                                                 \
    cudaError_t e__ = (expr);                                                          \
    if (e__ != cudaSuccess)                                                            \
      return fail(h, PXB_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(e__)); \
  } while (0)

#define PXB_REQUIRE_READY(h)                                                     \
  do {                                                                           \
    if (!(h)) return PXB_ERR_ARG;                                                \
    if (!(h)->arena) return fail(h, PXB_ERR_STATE, "arena not bound");           \
    if (!(h)->ham_set) return fail(h, PXB_ERR_STATE, "hamiltonian not set");     \
  } while (0)

inline cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// CTAs of the persistent local-energy kernels: all SMs, minus the ones the caller reserved for a
// concurrent side-stream kernel (pxb_reserve_sms; the serial comb plan needs one SM to itself)
inline int persist_sms(pxb_handle h) { return std::max(1, h->sm_count - h->reserved_sms); }

// brackets the launches of one stage with events when profiling is on
struct StageTimer {
  pxb_handle h;
  cudaStream_t st;
  int stage;
  cudaEvent_t a = nullptr, b = nullptr;
  static cudaEvent_t get(pxb_handle h) {
    cudaEvent_t e = nullptr;
    if (!h->evpool.empty()) {
      e = h->evpool.back();
      h->evpool.pop_back();
    } else if (cudaEventCreate(&e) != cudaSuccess) {
      e = nullptr;
    }
    return e;
  }
  StageTimer(pxb_handle h_, int stage_, cudaStream_t st_) : h(h_), st(st_), stage(stage_) {
    if (!h->prof) return;
    a = get(h);
    b = get(h);
    if (a) cudaEventRecord(a, st);
  }
  ~StageTimer() {
    if (!a || !b) return;
    cudaEventRecord(b, st);
    h->pending.push_back({stage, a, b});
  }
};

inline int grid_for(size_t n, int block = 256, int cap = 148 * 16) {
  size_t g = (n + block - 1) / block;
  if (g > (size_t)cap) g = cap;
  if (g < 1) g = 1;
  return (int)g;
}

CopyArgs copy_args(pxb_handle h) {
  CopyArgs c;
  c.phi = h->phi();
  c.theta = h->ptr<double>(A_THETA);
  c.e1b = h->ptr<double2>(A_E1B);
  c.weight = h->field<double>(PXB_F_WEIGHT);
  c.unscaled = h->field<double>(PXB_F_UNSCALED_WEIGHT);
  c.ot = h->field<double2>(PXB_F_OT);
  c.ehyb = h->field<double2>(PXB_F_HYBRID_ENERGY);
  c.eloc = h->field<double2>(PXB_F_ELOC);
  c.detR = h->field<double>(PXB_F_DETR);
  c.log_detR = h->field<double>(PXB_F_LOG_DETR);
  c.phase = h->field<double2>(PXB_F_PHASE);
  c.weloc = h->field<double2>(PXB_F_WALKER_ELOC);
  c.ottrue = h->ptr0<double2>(A_OT_TRUE);
  c.bpfac = h->ptr0<double2>(A_BPFAC);
  c.X = h->ptr<double2>(A_X);
  c.phi_old = h->nbp > 0 ? h->ptr<double>(A_PHI_OLD) : nullptr;
  c.fc = h->nbp > 0 ? h->ptr<double>(A_FC) : nullptr;
  c.fc_rows = (int)fc_rows(h->d, h->nbp);
  c.d = h->d;
  return c;
}

// ---- stage launchers ---------------------------------------------------------
template <int NMT>
int launch_greens(pxb_handle h, const GreensArgs& a, size_t smem, cudaStream_t st) {
  PXB_CUDA(h, cudaFuncSetAttribute(greens_kernel<NMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PXB_CUDA(h, cudaFuncSetAttribute(greens_kernel<NMT>, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   (int)cudaSharedmemCarveoutMaxShared));
  ++h->launches;
  greens_kernel<NMT><<<2 * h->d.Wp, GR_THREADS, smem, st>>>(a);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

template <int NMT>
int launch_theta(pxb_handle h, const double* phi, double* theta_out, cudaStream_t st);

template <int NMT>
int launch_greens2(pxb_handle h, const double* phi, cudaStream_t st) {
  const Dims& d = h->d;
  const int nmax = d.na > d.nb ? d.na : d.nb;
  const int nld = nmax | 1, nsq = nmax * nld;
  // 1. O_s = phi_s^T psi_s for all walkers; both spins in one launch (batch index = spin) when
  //    they have the same number of orbitals
  const bool both = d.na == d.nb && d.nb > 0;
  // complex orbitals: conj(psi)^T phi = [Re psi^T | -Im psi^T] [phi ; i phi] over a doubled k range
  const int km = kmul(d);
  const size_t kc = (size_t)km * d.KC;
  const double* pin = phi;
  if (km == 2) {
    ++h->launches;
    stack_rows_kernel<<<grid_for((size_t)d.WG * d.ne * d.KC * 16), 256, 0, st>>>(phi, h->ptr0<double>(A_PHI_STACK),
                                                                              (size_t)d.WG * d.ne, d.KC);
    PXB_CUDA(h, cudaGetLastError());
    pin = h->ptr0<double>(A_PHI_STACK);
  }
  for (int s = 0; s < (both ? 1 : 2); ++s) {
    const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
    if (ns == 0) continue;
    GemmArgs g;
    g.A = h->ptr<double>(A_PF) + (s ? (size_t)((d.na + 7) >> 3) * kc * 32 : 0);
    g.B = pin + (size_t)ioff * kc * 32;
    g.strideAz = (size_t)((d.na + 7) >> 3) * kc * 32;
    g.strideBz = (size_t)d.na * kc * 32;
    g.strideBO = (size_t)d.ne * kc * 32;
    g.strideBI = kc * 32;
    g.ntInner = ns;
    g.MTiles = (ns + 7) >> 3;
    g.NTiles = d.WG * ns;
    g.KS = (int)kc;
    EpiO epi{h->ptr<double2>(A_OB), ns, s, nld, nsq};
    ++h->launches;
    PXB_CUDA(h, (launch_gemm_tma<NMT, 4, 1, 6>(g, epi, both ? 2 : 1, h->sm_count, st)));
  }
  return launch_theta<NMT>(h, phi, h->ptr<double>(A_THETA), st);
}

// 2. inverse, slogdet, Theta, e1b: one warp per (walker, spin); Theta = OB^-1 phi^T -> theta_out
template <int NMT>
int launch_theta(pxb_handle h, const double* phi, double* theta_out, cudaStream_t st) {
  const Dims& d = h->d;
  const int nmax = d.na > d.nb ? d.na : d.nb;
  const int nld = nmax | 1, nsq = nmax * nld;
  ThetaArgs a;
  a.OB = h->ptr<double2>(A_OB);
  a.phi = phi;
  a.theta = theta_out;
  a.h1rot = h->ptr<double2>(A_H1ROT);
  a.slog = h->ptr<double>(A_SLOG);
  a.e1b_part = h->ptr<double2>(A_E1BP);
  a.d = d;
  a.nld = nld;
  a.nsq = nsq;
  const size_t spw = theta_smem_per_warp(nmax);
  const size_t smem = spw * TH_WARPS;
  if (smem > (size_t)h->max_smem_optin) return 1;
  PXB_CUDA(h, cudaFuncSetAttribute(theta_kernel<NMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ++h->launches;
  theta_kernel<NMT><<<(2 * d.Wp + TH_WARPS - 1) / TH_WARPS, TH_WARPS * 32, smem, st>>>(a, (int)spw);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

// phi = Q R in place (R_ii > 0), log det R per (walker, spin) -> A_QRLD.  Up to 32 orbitals per spin:
// CholeskyQR2 on DMMA (pxb_qr.cuh), then the Gram-Schmidt kernel for the items it marked; else
// Gram-Schmidt for everything.
template <int NMT>
static int launch_cholqr(pxb_handle h, double* phi, cudaStream_t st) {
  const Dims& d = h->d;
  const int nmax = d.na > d.nb ? d.na : d.nb;
  CholQrArgs a;
  a.phi = phi;
  a.logdet = h->ptr<double>(A_QRLD);
  a.need_mgs = h->ptr<int>(A_QRMASK);
  a.d = d;
  a.lda = nmax | 1;
  const size_t spw = cholqr_smem_per_warp(nmax);
  const size_t smem = spw * TH_WARPS;
  if (smem > (size_t)h->max_smem_optin) return 1;
  PXB_CUDA(h, cudaFuncSetAttribute(cholqr_kernel<NMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ++h->launches;
  cholqr_kernel<NMT><<<(2 * d.Wp + TH_WARPS - 1) / TH_WARPS, TH_WARPS * 32, smem, st>>>(a, (int)spw);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

static int run_qr(pxb_handle h, double* phi, cudaStream_t st) {
  const Dims& d = h->d;
  const int nmt = ((d.na > d.nb ? d.na : d.nb) + 7) >> 3;
  int rc = 1;
  if (h->qr_cholesky) {
    if (nmt <= 1) rc = launch_cholqr<1>(h, phi, st);
    else if (nmt <= 2) rc = launch_cholqr<2>(h, phi, st);
    else if (nmt <= 3) rc = launch_cholqr<3>(h, phi, st);
    else if (nmt <= 4) rc = launch_cholqr<4>(h, phi, st);
  }
  if (rc < 0) return rc;
  QrArgs a;
  a.phi = phi;
  a.logdet = h->ptr<double>(A_QRLD);
  a.d = d;
  a.mask = rc == PXB_OK ? h->ptr<int>(A_QRMASK) : nullptr;
  const size_t smem = qr_smem_bytes(d);
  if (smem > (size_t)h->max_smem_optin) return fail(h, PXB_ERR_ARG, "qr: problem too large for shared memory");
  ++h->launches;
  if (smem > (size_t)56 * 1024) {  // at most 3 CTAs per SM: more warps per problem instead
    PXB_CUDA(h, cudaFuncSetAttribute(qr_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qr_kernel<512><<<2 * d.Wp, 512, smem, st>>>(a);
  } else {
    PXB_CUDA(h, cudaFuncSetAttribute(qr_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    qr_kernel<128><<<2 * d.Wp, 128, smem, st>>>(a);
  }
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int run_greens(pxb_handle h, const double* phi, bool want_theta, double2* ovlp_out, bool with_e1b,
               cudaStream_t st) {
  StageTimer timer__(h, PXB_STAGE_GREENS, st);
  const Dims& d = h->d;
  GreensArgs a;
  a.phi = phi;
  a.theta = h->ptr<double>(A_THETA);
  a.PF = h->ptr<double>(A_PF);
  a.h1rot = h->ptr<double2>(A_H1ROT);
  a.slog = h->ptr<double>(A_SLOG);
  a.e1b_part = h->ptr<double2>(A_E1BP);
  a.d = d;
  a.want_theta = want_theta ? 1 : 0;
  const size_t smem = greens_smem_bytes(d);
  if (smem > (size_t)h->max_smem_optin) return fail(h, PXB_ERR_ARG, "greens: problem too large for shared memory");
  const int nmt = ((d.na > d.nb ? d.na : d.nb) + 7) >> 3;
  int rc = 1;
  if (h->greens_split && want_theta) {
    if (nmt <= 1) rc = launch_greens2<1>(h, phi, st);
    else if (nmt <= 2) rc = launch_greens2<2>(h, phi, st);
    else if (nmt <= 3) rc = launch_greens2<3>(h, phi, st);
    else if (nmt <= 4) rc = launch_greens2<4>(h, phi, st);
    else if (nmt <= 5) rc = launch_greens2<5>(h, phi, st);
    else if (nmt <= 6) rc = launch_greens2<6>(h, phi, st);
    else if (nmt <= 8) rc = launch_greens2<8>(h, phi, st);
  }
  if (rc != 1) {
    if (rc) return rc;
  } else if (kmul(d) == 2) {
    return fail(h, PXB_ERR_UNSUPPORTED, "complex trial orbitals: shape outside the split Green's function path");
  } else if (nmt <= 1) rc = launch_greens<1>(h, a, smem, st);
  else if (nmt <= 2) rc = launch_greens<2>(h, a, smem, st);
  else if (nmt <= 3) rc = launch_greens<3>(h, a, smem, st);
  else if (nmt <= 4) rc = launch_greens<4>(h, a, smem, st);
  else if (nmt <= 5) rc = launch_greens<5>(h, a, smem, st);
  else if (nmt <= 6) rc = launch_greens<6>(h, a, smem, st);
  else if (nmt <= 8) rc = launch_greens<8>(h, a, smem, st);
  else return fail(h, PXB_ERR_ARG, "greens: more than 64 occupied orbitals per spin");
  if (rc) return rc;
  ++h->launches;
  greens_combine_kernel<<<(d.Wp + 255) / 256, 256, 0, st>>>(
      a.slog, a.e1b_part, ovlp_out, (want_theta || with_e1b) ? h->ptr<double2>(A_E1B) : nullptr, d.Wp);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

// X_s[w][n] = sum_{i,p} R_s[(i,p), n] Theta_s[w][i][p]
int run_force_bias_gemm(pxb_handle h, cudaStream_t st) {
  StageTimer timer__(h, PXB_STAGE_XGEMM, st);
  const Dims& d = h->d;
  const bool both = d.na == d.nb && d.nb > 0;   // one launch, batch index = spin
  // complex Cholesky vectors: [Re R | Im R]^T [Theta ; i Theta], k range doubled per orbital row
  const int km = kmul(d);
  const size_t kc = (size_t)km * d.KC;
  const double* tin = h->ptr<double>(A_THETA);
  if (km == 2) {
    ++h->launches;
    stack_rows_kernel<<<grid_for((size_t)d.WG * d.ne * d.KC * 16), 256, 0, st>>>(tin, h->ptr0<double>(A_THETA_STACK),
                                                                              (size_t)d.WG * d.ne, d.KC);
    PXB_CUDA(h, cudaGetLastError());
    tin = h->ptr0<double>(A_THETA_STACK);
  }
  for (int s = 0; s < (both ? 1 : 2); ++s) {
    const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
    EpiX epi{h->ptr<double>(A_X) + (size_t)s * d.Wp * d.Np * 2, d.Wp, d.Np};
    if (ns == 0) {
      PXB_CUDA(h, cudaMemsetAsync(epi.X, 0, (size_t)d.Wp * d.Np * 16, st));
      continue;
    }
    GemmArgs g;
    g.A = h->ptr<double>(A_RF) + rf_spin_base(d, s);
    g.B = tin + (size_t)ioff * kc * 32;
    g.strideAz = rf_spin_base(d, 1);
    g.strideBz = (size_t)d.na * kc * 32;
    g.strideBO = (size_t)d.ne * kc * 32;
    g.strideBI = 0;
    g.ntInner = 1;
    g.MTiles = d.XG;
    g.NTiles = d.WG;
    g.KS = ns * (int)kc;
    ++h->launches;
    // 16 x 8 tile blocks: 4 * WG/8 work units keep the 148 persistent CTAs balanced (XG is only ~63)
    PXB_CUDA(h, (launch_gemm_tma<4, 4, 4, 2>(g, epi, both ? 2 : 1, persist_sms(h), st)));
  }
  return PXB_OK;
}

int run_vhs_gemm(pxb_handle h, cudaStream_t st) {
  StageTimer timer__(h, PXB_STAGE_VHS, st);
  const Dims& d = h->d;
  GemmArgs g;
  // complex Cholesky vectors: [Re L | Im L] [x ; i x] over a doubled k range
  const int km = kmul(d);
  const double* xin = h->ptr<double>(A_XF);
  if (km == 2) {
    ++h->launches;
    stack_rows_kernel<<<grid_for((size_t)d.WG * d.NKC * 16), 256, 0, st>>>(xin, h->ptr<double>(A_XF2), (size_t)d.WG,
                                                                        d.NKC);
    PXB_CUDA(h, cudaGetLastError());
    xin = h->ptr<double>(A_XF2);
  }
  g.A = h->ptr<double>(A_LF);
  g.B = xin;
  g.strideAz = g.strideBz = 0;
  g.strideBO = (size_t)km * d.NKC * 32;
  g.strideBI = 0;
  g.ntInner = 1;
  g.MTiles = h->vhs_sym ? h->rtu : d.RT;
  g.NTiles = d.WG;
  g.KS = km * d.NKC;
  EpiVHS epi{h->ptr<double>(A_VF), d.KC, d.MT, d.sqrt_dt, vf_walker(d),
             h->vhs_sym ? h->ptr<int>(A_RTMAP) : nullptr};
  ++h->launches;
  // CTA tile 16 x 16 tiles; 16 x 8 when the finer tiles fill the last wave of the persistent grid so much
  // better that it outweighs their lower reuse of L (1024 walkers per GPU: 11 half-size rounds
  // instead of 6 full ones)
  const int mb = (g.MTiles + 15) / 16, sms = h->sm_count;
  const int r16 = (mb * ((d.WG + 15) / 16) + sms - 1) / sms, r8 = (mb * ((d.WG + 7) / 8) + sms - 1) / sms;
  if (0.52 * r8 < 0.985 * r16) PXB_CUDA(h, (launch_gemm_tma<4, 4, 4, 2>(g, epi, 1, h->sm_count, st)));
  else PXB_CUDA(h, (launch_gemm_tma<4, 8, 4, 2>(g, epi, 1, h->sm_count, st)));
  return PXB_OK;
}

int run_one_body(pxb_handle h, const double* in, double* out, const int* active, cudaStream_t st,
                 bool transposed = false) {
  StageTimer timer__(h, PXB_STAGE_ONE_BODY, st);
  const Dims& d = h->d;
  // complex BH1: [Re BH1 | Im BH1] [phi ; i phi] -- the same real GEMM over a doubled k range
  const bool cplx = (d.flags & FLAG_COMPLEX_ONE_BODY) != 0;
  const int kc = cplx ? 2 * d.KC : d.KC;
  if (cplx) {
    if (transposed) return fail(h, PXB_ERR_UNSUPPORTED, "back propagation with a complex one-body propagator");
    ++h->launches;
    stack_rows_kernel<<<grid_for((size_t)d.WG * d.ne * d.KC * 16), 256, 0, st>>>(in, h->ptr<double>(A_PHI_STACK),
                                                                              (size_t)d.WG * d.ne, d.KC);
    PXB_CUDA(h, cudaGetLastError());
    in = h->ptr<double>(A_PHI_STACK);
  }
  const bool both = d.na == d.nb && d.nb > 0;   // one launch, batch index = spin
  for (int s = 0; s < (both ? 1 : 2); ++s) {
    const int ns = s ? d.nb : d.na, ioff = s ? d.na : 0;
    if (ns == 0) continue;
    GemmArgs g;
    g.A = (cplx ? h->ptr<double>(A_BF2) : h->ptr<double>(transposed ? A_BFT : A_BF)) + (size_t)s * d.MT * kc * 32;
    g.B = in + (size_t)ioff * kc * 32;
    g.strideAz = (size_t)d.MT * kc * 32;
    g.strideBz = (size_t)d.na * kc * 32;
    g.strideBO = (size_t)d.ne * kc * 32;
    g.strideBI = (size_t)kc * 32;
    g.ntInner = ns;
    g.MTiles = d.MT;
    g.NTiles = d.WG * ns;
    g.KS = kc;
    EpiOF epi{out, active, d.ne, d.KC, ioff, ns};
    ++h->launches;
    if (d.MT % 7 == 0) {
      PXB_CUDA(h, (launch_gemm_tma<7, 4, 2, 4>(g, epi, both ? 2 : 1, h->sm_count, st)));
    } else {
      PXB_CUDA(h, (launch_gemm_tma<4, 8, 4, 2>(g, epi, both ? 2 : 1, h->sm_count, st)));
    }
  }
  return PXB_OK;
}

template <int WMT, int NT, int NWARPS, int MINB>
int launch_taylor(pxb_handle h, TaylorArgs& a, int nwarps, cudaStream_t st) {
  const Dims& d = h->d;
  const size_t smem = (size_t)d.KC * NT * 32 * sizeof(double);
  if (smem > (size_t)h->max_smem_optin) return fail(h, PXB_ERR_ARG, "taylor: tile too large for shared memory");
  auto kern = taylor_kernel<WMT, NT, NWARPS, MINB>;
  PXB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  PXB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                   (int)cudaSharedmemCarveoutMaxShared));
  ++h->launches;
  kern<<<d.W * a.nchunks, nwarps * 32, smem, st>>>(a);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

template <int WMX, int WNX, int NG>
int launch_taylor2(pxb_handle h, const Taylor2Args& a, size_t smem, int grid, cudaStream_t st) {
  auto kern = taylor2_kernel<WMX, WNX, NG>;
  PXB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ++h->launches;
  kern<<<grid, T2Cfg<NG>::threads, smem, st>>>(a);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

// returns 1 if the shape is not covered by the persistent kernel (caller falls back), else status
int run_taylor2(pxb_handle h, double* phi, const int* active, int ochunk, int nchunks, cudaStream_t st) {
  const Dims& d = h->d;
  const int NT = ochunk / 4;
  if (d.MT < 4 || d.MT > 28 || NT < 2 || NT > 12) return 1;
  Taylor2Args a;
  a.VF = h->ptr<double>(A_VF);
  a.phi = phi;
  a.active = active;
  a.d = d;
  a.ochunk = ochunk;
  a.nchunks = nchunks;
  a.NT = NT;
  a.dbg = 0;
#ifdef PXB_EXPERIMENTS
  {
    const char* e = getenv("PXB_TAYLOR_DBG");
    if (e) a.dbg = atoi(e);
  }
#endif
  // 4 column groups (16 consumer warps, 112 registers) unless the m-groups are too tall for that budget
  int NG = (NT >= 4 && (d.MT + 3) / 4 <= 4) ? 4 : 2;
#ifdef PXB_EXPERIMENTS
  {
    const char* e = getenv("PXB_TAYLOR_GROUPS");
    if (e && atoi(e) == 2) NG = 2;
  }
#endif
  // 4 m-groups and NG column groups, sizes differing by at most one, larger ones first
  const int base = d.MT / 4, rem = d.MT % 4;
  int msize[4];
  a.m_off[0] = 0;
  for (int g = 0; g < 4; ++g) {
    msize[g] = base + (g < rem ? 1 : 0);
    a.m_off[g + 1] = a.m_off[g] + msize[g];
  }
  const int nbase = NT / NG, nrem = NT % NG;
  int nsize[T2_MAXG] = {0, 0, 0, 0};
  a.n_off[0] = 0;
  for (int g = 0; g < NG; ++g) {
    nsize[g] = nbase + (g < nrem ? 1 : 0);
    a.n_off[g + 1] = a.n_off[g] + nsize[g];
  }
  for (int g = NG; g < T2_MAXG; ++g) a.n_off[g + 1] = a.n_off[NG];
  // m-group permutation per column group: the largest remaining m-group goes to the least loaded
  // sub-partition (column groups are already in descending size)
  int load[4] = {0, 0, 0, 0};
  for (int g = 0; g < T2_MAXG; ++g) {
    int order[4] = {0, 1, 2, 3};
    std::sort(order, order + 4, [&](int x, int y) { return load[x] != load[y] ? load[x] < load[y] : x < y; });
    for (int k = 0; k < 4; ++k) {
      a.mperm[g][order[k]] = k;  // m-groups are sorted by descending size
      if (g < NG) load[order[k]] += msize[k] * nsize[g];
    }
  }
  const int wmx = base + (rem ? 1 : 0), wnx = nbase + (nrem ? 1 : 0);
  // shared memory: two iterate buffers when a >= 3-deep ring still fits, else one
  const size_t budget = (size_t)h->max_smem_optin;
  a.nbuf = 0;
  int nbuf_max = 2;
#ifdef PXB_EXPERIMENTS
  {
    const char* e = getenv("PXB_TAYLOR_NBUF");
    if (e && atoi(e) == 1) nbuf_max = 1;
  }
#endif
  for (int nbuf = nbuf_max; nbuf >= 1 && a.nbuf == 0; --nbuf)
    for (int nstage = 12; nstage >= 3; --nstage)
      if (taylor2_smem_bytes(d, NT, nbuf, nstage) <= budget) {
        a.nbuf = nbuf;
        a.nstage = nstage;
        break;
      }
  if (a.nbuf == 0) return 1;
  const size_t smem = taylor2_smem_bytes(d, NT, a.nbuf, a.nstage);
  const int grid = std::min(d.W * nchunks, h->sm_count);
#define PXB_T2(WM_, WN_, NG_) \
  if (wmx == WM_ && wnx == WN_ && NG == NG_) return launch_taylor2<WM_, WN_, NG_>(h, a, smem, grid, st);
  PXB_T2(1, 1, 4) PXB_T2(1, 2, 4) PXB_T2(1, 3, 4)
  PXB_T2(2, 1, 4) PXB_T2(2, 2, 4) PXB_T2(2, 3, 4)
  PXB_T2(3, 1, 4) PXB_T2(3, 2, 4) PXB_T2(3, 3, 4)
  PXB_T2(4, 1, 4) PXB_T2(4, 2, 4) PXB_T2(4, 3, 4)
  PXB_T2(1, 1, 2) PXB_T2(1, 2, 2) PXB_T2(2, 1, 2) PXB_T2(2, 2, 2) PXB_T2(3, 1, 2) PXB_T2(3, 2, 2)
  PXB_T2(4, 1, 2) PXB_T2(4, 2, 2) PXB_T2(4, 6, 2)
  PXB_T2(5, 1, 2) PXB_T2(5, 2, 2) PXB_T2(5, 3, 2) PXB_T2(5, 4, 2) PXB_T2(5, 5, 2) PXB_T2(5, 6, 2)
  PXB_T2(6, 1, 2) PXB_T2(6, 2, 2) PXB_T2(6, 3, 2) PXB_T2(6, 4, 2) PXB_T2(6, 5, 2) PXB_T2(6, 6, 2)
  PXB_T2(7, 1, 2) PXB_T2(7, 2, 2) PXB_T2(7, 3, 2) PXB_T2(7, 4, 2) PXB_T2(7, 5, 2) PXB_T2(7, 6, 2)
#undef PXB_T2
  return 1;
}

template <int WMX, int WNX, int NG, int MG = 4>
int launch_taylor3(pxb_handle h, const Taylor3Args& a, size_t smem, int grid, cudaStream_t st) {
  auto kern = taylor3_kernel<WMX, WNX, NG, MG>;
  PXB_CUDA(h, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  ++h->launches;
  kern<<<grid, T3Cfg<NG, MG>::threads, smem, st>>>(a);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

// 3-product planar kernel (pxb_taylor3.cuh): shapes whose warp rectangle (ceil(MT/4) x ceil(NT8/NG)
// tile pairs x 3 accumulator sets) fits the consumer register budget (NG = 3: 160 registers, up to
// 4 x 2; NG = 2: 232 registers, up to 4 x 3) and whose tile buffers leave room for a >= 3-deep ring.
// Returns 1 if the shape is not covered.
int run_taylor3(pxb_handle h, double* phi, const int* active, cudaStream_t st) {
  const Dims& d = h->d;
  int nchunks = (d.ne + 47) / 48;
  const int ochunk = round_up((d.ne + nchunks - 1) / nchunks, 8);
  nchunks = (d.ne + ochunk - 1) / ochunk;
  const int NT8 = ochunk / 8;
  if (d.MT < 4 || d.MT > 16 || NT8 < 2 || NT8 > 6) return 1;
  Taylor3Args a;
  a.VF = h->ptr<double>(A_VF);
  a.phi = phi;
  a.active = active;
  a.d = d;
  a.ochunk = ochunk;
  a.nchunks = nchunks;
  a.NT8 = NT8;
  a.S = NT8 * 64 + 4;
  const int base = d.MT / 4, rem = d.MT % 4;
  const int wmx = base + (rem ? 1 : 0);
  // three column groups (12 consumer warps) when the columns split evenly and the rectangle fits
  // 160 registers, else two (8 warps at 232 registers); with at most 16 columns ONE group of 4 warps
  // and two CTAs per SM (two walkers in flight) instead of 6 DMMAs per k-step and warp
  int NG = (NT8 % 3 == 0 && wmx * (NT8 / 3) <= 8) ? 3 : 2;
  if (NT8 <= 2 && wmx * NT8 <= 6) NG = 1;
#ifdef PXB_EXPERIMENTS
  {
    const char* e = getenv("PXB_TAYLOR_GROUPS");
    if (e && atoi(e) == 2) NG = 2;
  }
#endif
  // 48 columns and an even number of m-tiles (c4: 14 x 6 tile pairs): six column groups of TWO warps,
  // 7 x 1 tile pairs per warp -- every sub-partition gets the same 21 tile pairs (22 : 20 with three
  // groups of four warps) and a group rendez-vous is between two warps
  bool six = NT8 == 6 && d.MT % 2 == 0 && d.MT / 2 >= 5 && d.MT / 2 <= 7;
#ifdef PXB_EXPERIMENTS
  {
    const char* e = getenv("PXB_TAYLOR_GROUPS");
    if (e && atoi(e) != 6) six = false;
  }
#endif
  if (six) {
    const int half = d.MT / 2;
    a.m_off[0] = 0;
    a.m_off[1] = half;
    a.m_off[2] = a.m_off[3] = a.m_off[4] = d.MT;
    for (int g = 0; g <= 6; ++g) a.n_off[g] = g;
    for (int g = 0; g < 6; ++g)
      for (int k = 0; k < 4; ++k) a.mperm[g][k] = k & 1;
    a.nstage = 0;
    a.nbuf = 0;
    a.dbg = 0;
#ifdef PXB_EXPERIMENTS
    {
      const char* e3 = getenv("PXB_T3_DBG");
      if (e3) a.dbg = atoi(e3);
    }
#endif
    for (int nbuf = 3; nbuf >= 2 && a.nbuf == 0; --nbuf)
      for (int nstage = 12; nstage >= (nbuf == 3 ? 4 : 3); --nstage)
        if (taylor3_smem_bytes(d, NT8, nbuf, nstage) <= (size_t)h->max_smem_optin) {
          a.nbuf = nbuf;
          a.nstage = nstage;
          break;
        }
    if (a.nbuf != 0) {
      const size_t smem6 = taylor3_smem_bytes(d, NT8, a.nbuf, a.nstage);
      const int grid6 = std::min(d.W * nchunks, h->sm_count);
      if (half == 7) return launch_taylor3<7, 1, 6, 2>(h, a, smem6, grid6, st);
      if (half == 6) return launch_taylor3<6, 1, 6, 2>(h, a, smem6, grid6, st);
      return launch_taylor3<5, 1, 6, 2>(h, a, smem6, grid6, st);
    }
  }
  const int ctas = NG == 1 ? 2 : 1;
  int msize[4];
  a.m_off[0] = 0;
  for (int g = 0; g < 4; ++g) {
    msize[g] = base + (g < rem ? 1 : 0);
    a.m_off[g + 1] = a.m_off[g] + msize[g];
  }
  const int nbase = NT8 / NG, nrem = NT8 % NG;
  int nsize[3] = {0, 0, 0};
  a.n_off[0] = 0;
  for (int g = 0; g < NG; ++g) {
    nsize[g] = nbase + (g < nrem ? 1 : 0);
    a.n_off[g + 1] = a.n_off[g] + nsize[g];
  }
  for (int g = NG; g < 3; ++g) a.n_off[g + 1] = a.n_off[NG];
  // m-group permutation per column group: the largest remaining m-group goes to the least loaded
  // sub-partition (column groups are in descending size)
  int load[4] = {0, 0, 0, 0};
  for (int g = 0; g < 3; ++g) {
    int order[4] = {0, 1, 2, 3};
    std::sort(order, order + 4, [&](int x, int y) { return load[x] != load[y] ? load[x] < load[y] : x < y; });
    for (int k = 0; k < 4; ++k) {
      a.mperm[g][order[k]] = k;
      if (g < NG) load[order[k]] += msize[k] * nsize[g];
    }
  }
  const int wnx = nbase + (nrem ? 1 : 0);
  if (wmx * wnx > (NG == 3 ? 8 : NG == 2 ? 12 : 6)) return 1;
  // shared memory: iterate buffer + phi tile (+ a second phi tile, fetched one item ahead, when a
  // >= 4-deep ring still fits beside it)
  a.nstage = 0;
  a.nbuf = 0;
  a.dbg = 0;
#ifdef PXB_EXPERIMENTS
  {
    const char* e3 = getenv("PXB_T3_DBG");
    if (e3) a.dbg = atoi(e3);
  }
#endif
  for (int nbuf = 3; nbuf >= 2 && a.nbuf == 0; --nbuf)
    for (int nstage = 12; nstage >= (nbuf == 3 ? 4 : 3); --nstage)
      if (taylor3_smem_bytes(d, NT8, nbuf, nstage) <= (size_t)h->max_smem_optin / ctas - (ctas - 1) * 1024) {
        a.nbuf = nbuf;
        a.nstage = nstage;
        break;
      }
  if (a.nbuf == 0) return 1;
  const size_t smem = taylor3_smem_bytes(d, NT8, a.nbuf, a.nstage);
  const int grid = std::min(d.W * nchunks, ctas * h->sm_count);
#define PXB_T3(WM_, WN_, NG_) \
  if (wmx == WM_ && wnx == WN_ && NG == NG_) return launch_taylor3<WM_, WN_, NG_>(h, a, smem, grid, st);
  PXB_T3(1, 1, 3) PXB_T3(1, 2, 3) PXB_T3(2, 1, 3) PXB_T3(2, 2, 3) PXB_T3(3, 1, 3) PXB_T3(3, 2, 3)
  PXB_T3(4, 1, 3) PXB_T3(4, 2, 3)
  PXB_T3(1, 1, 2) PXB_T3(1, 2, 2) PXB_T3(1, 3, 2) PXB_T3(2, 1, 2) PXB_T3(2, 2, 2) PXB_T3(2, 3, 2)
  PXB_T3(3, 1, 2) PXB_T3(3, 2, 2) PXB_T3(3, 3, 2) PXB_T3(4, 1, 2) PXB_T3(4, 2, 2) PXB_T3(4, 3, 2)
  PXB_T3(1, 2, 1) PXB_T3(2, 2, 1) PXB_T3(3, 2, 1)
#undef PXB_T3
  return 1;
}

int run_taylor(pxb_handle h, double* phi, const int* active, cudaStream_t st) {
  StageTimer timer__(h, PXB_STAGE_TAYLOR, st);
  const Dims& d = h->d;
  TaylorArgs a;
  a.VF = h->ptr<double>(A_VF);
  a.phi = phi;
  a.active = active;
  a.d = d;
  // orbitals per CTA: at most 48 (12 n-tiles), balanced over chunks, multiple of 4
  int nchunks = (d.ne + 47) / 48;
  int ochunk = round_up((d.ne + nchunks - 1) / nchunks, 4);
  nchunks = (d.ne + ochunk - 1) / ochunk;
  a.nchunks = nchunks;
  const int NT = ochunk / 4;
  if (h->taylor_tma && h->taylor_3m) {
    const int rc3 = run_taylor3(h, phi, active, st);
    if (rc3 != 1) return rc3;
  }
  if (h->taylor_tma) {
    const int rc2 = run_taylor2(h, phi, active, ochunk, nchunks, st);
    if (rc2 != 1) return rc2;
  }
  // m-tiles per warp: 2 when that fits in 8 warps, else as many as needed
  int wmt = (d.MT + 7) / 8;
  if (wmt < 2 && d.MT >= 2) wmt = 2;
  if (wmt > 4) return fail(h, PXB_ERR_ARG, "taylor: nbasis > 256 not supported in this version");
  const int nwarps = (d.MT + wmt - 1) / wmt;
  // n-tiles are a compile-time constant: the chunk is padded with zero orbitals up to it
#define PXB_TAYLOR_CASE(W_, N_, NW_, MB_)                   \
  if (wmt == W_ && NT <= N_ && nwarps <= NW_) {            \
    a.ochunk = ochunk;                                     \
    return launch_taylor<W_, N_, NW_, MB_>(h, a, nwarps, st); \
  }
  PXB_TAYLOR_CASE(1, 2, 8, 2)
  PXB_TAYLOR_CASE(1, 4, 8, 2)
  PXB_TAYLOR_CASE(1, 8, 8, 2)
  PXB_TAYLOR_CASE(1, 12, 8, 2)
  PXB_TAYLOR_CASE(2, 2, 8, 2)
  PXB_TAYLOR_CASE(2, 4, 8, 2)
  PXB_TAYLOR_CASE(2, 6, 8, 2)
  PXB_TAYLOR_CASE(2, 8, 8, 2)
  PXB_TAYLOR_CASE(2, 10, 8, 2)
  PXB_TAYLOR_CASE(2, 11, 7, 2)
  PXB_TAYLOR_CASE(2, 12, 8, 1)
  PXB_TAYLOR_CASE(3, 10, 8, 1)
  PXB_TAYLOR_CASE(3, 12, 8, 1)
  PXB_TAYLOR_CASE(4, 10, 8, 1)
  PXB_TAYLOR_CASE(4, 12, 8, 1)
#undef PXB_TAYLOR_CASE
  return fail(h, PXB_ERR_ARG, "taylor: no kernel instance for this shape");
}

int run_exchange_eri(pxb_handle h, cudaStream_t st) {
  const Dims& d = h->d;
  EriArgs a;
  a.theta = h->ptr<double>(A_THETA);
  a.d = d;
  a.nslot = h->eri_nslot;
  const int nrb = std::max(eri_rowblocks(d, 0), eri_rowblocks(d, 1));
  const int nwb = (d.WG + EQ_TN - 1) / EQ_TN;
  const int nitems = nrb * 2 * nwb;
  // work counter of the dynamic item scheduler: lives behind the partial sums, zero from the arena
  // memset and put back to zero by the reduce kernel that follows every exchange launch
  double2* part_r = h->ptr<double2>(A_EPART);
  int* counter = reinterpret_cast<int*>(part_r + (size_t)2 * a.nslot * d.Wp);
  const size_t smem = eri_smem_bytes();
  PXB_CUDA(h, cudaFuncSetAttribute(exx_eri_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  // walker blocks per super-block: Theta of a super-block <= 64 MB, super-blocks of equal size
  const size_t theta_wb = (size_t)EQ_TN * d.ne * d.KC * 32 * 8;
  const int sbmax = (int)std::max<size_t>(1, ((size_t)64 << 20) / theta_wb);
  const int nsb = (nwb + sbmax - 1) / sbmax;
  const int SB = (nwb + nsb - 1) / nsb;
  // complex Cholesky vectors: K = Kr + i Ki, two passes of the same real quadratic form
  const bool cc = kmul(d) == 2;
  for (int pass = 0; pass < (cc ? 2 : 1); ++pass) {
    a.KF[0] = h->ptr<double>(pass ? A_KF0I : A_KF0);
    a.KF[1] = h->kf_shared[h->det] ? a.KF[0] : h->ptr<double>(pass ? A_KF1I : A_KF1);
    a.part = pass ? h->ptr<double2>(A_EPART2) : part_r;
    if (pass) PXB_CUDA(h, cudaMemsetAsync(counter, 0, 4, st));
    ++h->launches;
    exx_eri_kernel<<<std::min(nitems, persist_sms(h)), gemm_tma_threads<EQ_CWM * EQ_CWN>(), smem, st>>>(a, nitems, 2 * nrb, SB, counter);
    PXB_CUDA(h, cudaGetLastError());
  }
  ++h->launches;
  exx_eri_reduce_kernel<<<(2 * d.Wp + 255) / 256, 256, 0, st>>>(part_r, cc ? h->ptr<double2>(A_EPART2) : nullptr,
                                                              h->ptr<double2>(A_EXX), d, a.nslot, counter);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int run_exchange(pxb_handle h, cudaStream_t st) {
  StageTimer timer__(h, PXB_STAGE_EXCHANGE, st);
  const Dims& d = h->d;
  if (h->exx_eri) return run_exchange_eri(h, st);
  if (kmul(d) == 2)
    return fail(h, PXB_ERR_UNSUPPORTED, "complex Cholesky vectors need the ERI form of the exchange (exchange_mode)");
  ExArgs a;
  a.RF = h->ptr<double>(A_RF);
  a.theta = h->ptr<double>(A_THETA);
  a.exx = h->ptr<double>(A_EXX);
  a.d = d;
  const int nmax = d.na > d.nb ? d.na : d.nb;
  if ((nmax + 3) / 4 > EX_MAX_BLOCKS) return fail(h, PXB_ERR_ARG, "exchange: more than 64 occupied orbitals per spin");
  const size_t bbytes = (size_t)nmax * d.KC * 32 * 8;
  const size_t tail = exchange_tail_bytes();
  const int grid = std::min(2 * d.WG, persist_sms(h));
  if (bbytes + tail <= (size_t)h->max_smem_optin) {
    a.smem_b_doubles = (int)(bbytes / 8);
    PXB_CUDA(h, cudaFuncSetAttribute(exchange_kernel<4, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                     (int)(bbytes + tail)));
    ++h->launches;
    exchange_kernel<4, true><<<grid, EX_WARPS * 32, bbytes + tail, st>>>(a);
  } else {
    a.smem_b_doubles = 0;
    ++h->launches;
    exchange_kernel<4, false><<<grid, EX_WARPS * 32, tail, st>>>(a);
  }
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

// Green's function stage for every determinant of the trial: Theta_i, e1b_i and the determinant
// overlaps; ovlp_out receives the trial overlap sum_i conj(c_i) <psi_i|phi> (multi_det.py:198-231)
int run_greens_all(pxb_handle h, const double* phi, double2* ovlp_out, cudaStream_t st) {
  if (h->ndets == 1) return run_greens(h, phi, true, ovlp_out, true, st);
  int rc = PXB_OK;
  for (int i = 0; i < h->ndets && rc == PXB_OK; ++i) {
    h->det = i;
    rc = run_greens(h, phi, true, h->ptr<double2>(A_OVLP_DET), true, st);
  }
  h->det = 0;
  if (rc) return rc;
  const Dims& d = h->d;
  ++h->launches;
  md_overlap_kernel<<<(d.Wp + 255) / 256, 256, 0, st>>>(h->ptr0<double2>(A_COEFF), h->ptr0<double2>(A_OVLP_DET),
                                                        h->reg[A_OVLP_DET].stride / 16, h->ndets, ovlp_out, d.Wp);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

// Theta, overlap and e1b of the CURRENT walkers (recomputed only when stale)
int ensure_theta(pxb_handle h, cudaStream_t st) {
  if (h->theta_valid) return PXB_OK;
  int rc = run_greens_all(h, h->phi(), h->ptr0<double2>(A_OVLP_OLD), st);
  if (rc) return rc;
  h->theta_valid = true;
  h->x_valid = false;
  return PXB_OK;
}

// X_i = R_i^T Theta_i for every determinant; with several, also the weighted average that the
// force bias uses (propagation/generic.py:154-157) as a combined X for field_kernel
int ensure_x(pxb_handle h, cudaStream_t st) {
  if (h->x_valid) return PXB_OK;
  int rc = PXB_OK;
  for (int i = 0; i < h->ndets && rc == PXB_OK; ++i) {
    h->det = i;
    rc = run_force_bias_gemm(h, st);
  }
  h->det = 0;
  if (rc) return rc;
  if (h->ndets > 1) {
    const Dims& d = h->d;
    ++h->launches;
    md_x_kernel<<<dim3((d.Np + 255) / 256, d.Wp), 256, 0, st>>>(
        h->ptr0<double2>(A_COEFF), h->ptr0<double2>(A_OVLP_DET), h->reg[A_OVLP_DET].stride / 16,
        h->ptr0<double2>(A_X), h->reg[A_X].stride / 16, h->ndets, h->ptr0<double2>(A_XC), d);
    PXB_CUDA(h, cudaGetLastError());
  }
  h->x_valid = true;
  return PXB_OK;
}

// exchange + energy assembly of every determinant into eloc_det (or straight into the walker
// field for a single determinant when to_field is set)
int run_det_energies(pxb_handle h, bool to_field, cudaStream_t st) {
  const Dims& d = h->d;
  int rc = PXB_OK;
  for (int i = 0; i < h->ndets && rc == PXB_OK; ++i) {
    h->det = i;
    if ((rc = run_exchange(h, st))) break;
    EnergyArgs e;
    e.X = h->ptr<double2>(A_X);
    e.exx = h->ptr<double2>(A_EXX);
    e.e1b = h->ptr<double2>(A_E1B);
    e.eloc = (to_field && h->ndets == 1) ? h->field<double2>(PXB_F_ELOC) : h->ptr<double2>(A_ELOC_DET);
    e.d = d;
    StageTimer timer__(h, PXB_STAGE_ENERGY, st);
    ++h->launches;
    energy_kernel<<<(d.W + 7) / 8, 256, 0, st>>>(e);
    if (cudaGetLastError() != cudaSuccess) rc = fail(h, PXB_ERR_CUDA, "energy_kernel launch failed");
  }
  h->det = 0;
  return rc;
}

// sum_i w_i E_i / sum_i w_i with the CURRENT determinant overlaps (estimators/mixed.py:439-448)
int run_md_energy(pxb_handle h, double2* out, int only_total, cudaStream_t st) {
  const Dims& d = h->d;
  ++h->launches;
  md_energy_kernel<<<(d.W + 255) / 256, 256, 0, st>>>(h->ptr0<double2>(A_COEFF), h->ptr0<double2>(A_OVLP_DET),
                                                      h->reg[A_OVLP_DET].stride / 16, h->ptr0<double2>(A_ELOC_DET),
                                                      h->reg[A_ELOC_DET].stride / 16, h->ndets, out, only_total, d.W);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int set_step_params(pxb_handle h, const StepParams& v, int mask, cudaStream_t st) {
  ++h->launches;
  set_step_params_kernel<<<1, 1, 0, st>>>(h->ptr<StepParams>(A_STEP_PARAMS), v, mask);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

}  // namespace

// =============================================================================
extern "C" {

int pxb_abi_version(void) { return PXB_ABI_VERSION; }

long long pxb_launch_count(pxb_handle h) { return h ? h->launches : -1; }

int pxb_exchange_mode(pxb_handle h) {
  if (!h) return PXB_ERR_ARG;
  return h->exx_eri ? PXB_EXCHANGE_ERI : PXB_EXCHANGE_CHOLESKY;
}

int pxb_vhs_symmetric(pxb_handle h) {
  if (!h) return PXB_ERR_ARG;
  return h->vhs_sym ? 1 : 0;
}

int pxb_profile(pxb_handle h, int enable) {
  if (!h) return PXB_ERR_ARG;
  h->prof = enable != 0;
  return PXB_OK;
}

int pxb_stage_times(pxb_handle h, double* ms, long long* calls, int n, int reset) {
  if (!h) return PXB_ERR_ARG;
  for (auto& p : h->pending) {
    float t = 0.f;
    PXB_CUDA(h, cudaEventSynchronize(p.b));
    PXB_CUDA(h, cudaEventElapsedTime(&t, p.a, p.b));
    h->stage_ms[p.stage] += t;
    h->stage_calls[p.stage] += 1;
    h->evpool.push_back(p.a);
    h->evpool.push_back(p.b);
  }
  h->pending.clear();
  for (int i = 0; i < n && i < PXB_STAGE_COUNT; ++i) {
    if (ms) ms[i] = h->stage_ms[i];
    if (calls) calls[i] = h->stage_calls[i];
  }
  if (reset)
    for (int i = 0; i < PXB_STAGE_COUNT; ++i) {
      h->stage_ms[i] = 0.0;
      h->stage_calls[i] = 0;
    }
  return PXB_OK;
}

const char* pxb_last_error(pxb_handle h) { return h ? h->err.c_str() : "null handle"; }

int pxb_create(pxb_handle* out, const pxb_config* cfg) {
  if (!out || !cfg) return PXB_ERR_ARG;
  *out = nullptr;
  if (cfg->nbasis < 1 || cfg->nup < 1 || cfg->ndown < 0 || cfg->nchol < 1 || cfg->nwalkers < 1 ||
      cfg->nup > cfg->nbasis || cfg->ndown > cfg->nbasis || cfg->dt <= 0.0 || cfg->exp_order < 1)
    return PXB_ERR_ARG;
  pxb_context* h = new (std::nothrow) pxb_context();
  if (!h) return PXB_ERR_ARG;
  h->cfg = *cfg;
  h->ndets = cfg->ndets > 1 ? cfg->ndets : 1;
  if (h->ndets > PXB_MAX_DETS || (h->ndets > 1 && cfg->nbp > 0)) {
    delete h;
    return PXB_ERR_ARG;  // at most PXB_MAX_DETS determinants; no back propagation with several
  }
#ifdef PXB_EXPERIMENTS  // kernel-variant switches of development builds only (not in the product library)
  {
    const char* gr = getenv("PXB_GREENS");
    if (gr && strcmp(gr, "fused") == 0) h->greens_split = false;
    const char* v = getenv("PXB_VHS");
    if (v && strcmp(v, "full") == 0) h->vhs_sym_allowed = false;
    const char* t = getenv("PXB_TAYLOR");
    if (t && strcmp(t, "direct") == 0) h->taylor_tma = false;
    if (t && strcmp(t, "4m") == 0) h->taylor_3m = false;
    const char* q = getenv("PXB_QR");
    if (q && strcmp(q, "mgs") == 0) h->qr_cholesky = false;
  }
#endif
  Dims& d = h->d;
  d.M = cfg->nbasis;
  d.na = cfg->nup;
  d.nb = cfg->ndown;
  d.ne = d.na + d.nb;
  d.N = cfg->nchol;
  d.W = cfg->nwalkers;
  d.Wtot = cfg->total_walkers > 0 ? cfg->total_walkers : cfg->nwalkers;
  d.Mp = round_up(d.M, 4);
  d.KC = d.Mp / 4;
  d.M8 = round_up(d.M, 8);
  d.MT = d.M8 / 8;
  d.Wp = round_up(d.W, 4);
  d.WG = d.Wp / 4;
  d.Np = round_up(d.N, 8);
  d.XG = d.Np / 8;
  d.NKC = d.Np / 4;
  d.RT = ((d.M + 1) / 2) * d.KC;
  d.exp_order = cfg->exp_order;
  d.flags = cfg->flags;
  if (((d.flags & FLAG_LOCAL_ENERGY_WEIGHT) && (d.flags & FLAG_FREE_PROJECTION)) ||
      ((d.flags & (FLAG_COMPLEX_ONE_BODY | FLAG_COMPLEX_CHOLESKY)) && cfg->nbp > 0) ||
      ((d.flags & FLAG_COMPLEX_CHOLESKY) && cfg->exchange_mode == PXB_EXCHANGE_CHOLESKY)) {
    delete h;
    return PXB_ERR_ARG;
  }
  if (d.flags & FLAG_FREE_PROJECTION) d.flags |= FLAG_NO_FORCE_BIAS;  // continuous.py:30-33
  d.dt = cfg->dt;
  d.sqrt_dt = sqrt(cfg->dt);
  d.ebound = sqrt(2.0 / cfg->dt);
  d.ecore = 0.0;

  cudaError_t e = cudaSetDevice(cfg->device);
  if (e == cudaSuccess) {
    cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, cfg->device);
    cudaDeviceGetAttribute(&h->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, cfg->device);
  } else {
    // no device (e.g. symbol-load check on a CPU box): keep defaults, compute calls will fail loudly
    cudaGetLastError();
    h->max_smem_optin = 227 * 1024;
  }

  {
    const size_t kbytes = (eri_kf_doubles(d, 0) + eri_kf_doubles(d, 1)) * 8 * (size_t)h->ndets * kmul(d);
    if (cfg->exchange_mode == PXB_EXCHANGE_ERI)
      h->exx_eri = true;
    else if (cfg->exchange_mode == PXB_EXCHANGE_CHOLESKY)
      h->exx_eri = false;
    else if (cfg->exchange_mode == PXB_EXCHANGE_AUTO)
      h->exx_eri = kmul(d) == 2 || kbytes <= ((size_t)16 << 30);
    else {
      delete h;
      return PXB_ERR_ARG;
    }
    h->eri_nslot = std::max(eri_rowblocks(d, 0), eri_rowblocks(d, 1)) * EQ_CWM;
  }

  size_t off = 0;
  auto add = [&](int id, size_t bytes) {
    h->reg[id].off = off;
    h->reg[id].bytes = bytes;
    off += (bytes + 255) / 256 * 256;
  };
  // one copy per determinant of the trial
  const int D = h->ndets;
  auto addd = [&](int id, size_t bytes) {
    h->reg[id].off = off;
    h->reg[id].bytes = bytes;
    h->reg[id].stride = (bytes + 255) / 256 * 256;
    off += h->reg[id].stride * (size_t)D;
  };
  const size_t W = d.Wp, Wt = (size_t)d.Wtot;
  add(A_LF, lf_size(d) * 8);
  addd(A_RF, rf_size(d) * 8);
  add(A_BF, bf_size(d) * 8);
  addd(A_PSIT, (size_t)d.ne * d.Mp * 8);
  addd(A_H1ROT, (size_t)d.ne * d.Mp * 16);
  add(A_VBAR, (size_t)d.Np * 16);
  add(A_PHI_A, of_size(d) * 8);
  add(A_PHI_B, of_size(d) * 8);
  addd(A_THETA, of_size(d) * 8);
  addd(A_X, (size_t)2 * W * d.Np * 16);
  add(A_XF, xf_size(d) * 8);
  add(A_VF, (size_t)W * vf_walker(d) * 8);
  addd(A_EXX, 2 * W * 16);
  addd(A_KF0, h->exx_eri ? eri_kf_doubles(d, 0) * 8 : 0);
  addd(A_KF1, h->exx_eri ? eri_kf_doubles(d, 1) * 8 : 0);
  add(A_RTMAP, (size_t)d.RT * 4);
  {
    const size_t nmax = d.na > d.nb ? d.na : d.nb;
    add(A_OB, W * 2 * nmax * (nmax | 1) * 16);
  }
  add(A_EPART, h->exx_eri ? (size_t)2 * h->eri_nslot * W * 16 + 16 : 0);  // + the item counter of exx_eri_kernel
  {  // complex Cholesky vectors: stacked copies of Theta and of the fields, imaginary part of K
    const bool cc = kmul(d) == 2;
    add(A_THETA_STACK, cc ? 2 * of_size(d) * 8 : 0);
    add(A_XF2, cc ? 2 * xf_size(d) * 8 : 0);
    addd(A_KF0I, cc && h->exx_eri ? eri_kf_doubles(d, 0) * 8 : 0);
    addd(A_KF1I, cc && h->exx_eri ? eri_kf_doubles(d, 1) * 8 : 0);
    add(A_EPART2, cc && h->exx_eri ? (size_t)2 * h->eri_nslot * W * 16 : 0);
  }
  addd(A_E1B, W * 16);
  add(A_OVLP_OLD, W * 16);
  add(A_ACTIVE, W * 4);
  add(A_GW, Wt * 8);
  add(A_GWS, Wt * 8);
  add(A_CPROBS, Wt * 8);
  add(A_FLAG, 256);
  addd(A_PF, (size_t)(((d.na + 7) >> 3) + ((d.nb + 7) >> 3)) * d.KC * 32 * 8 * kmul(d));
  add(A_SLOG, W * 8 * 8);
  add(A_E1BP, W * 2 * 16);
  add(A_QRLD, W * 2 * 8);
  add(A_QRMASK, W * 2 * 4);
  add(A_FIELD0 + PXB_F_WEIGHT, W * 8);
  add(A_FIELD0 + PXB_F_UNSCALED_WEIGHT, W * 8);
  add(A_FIELD0 + PXB_F_OT, W * 16);
  add(A_FIELD0 + PXB_F_HYBRID_ENERGY, W * 16);
  add(A_FIELD0 + PXB_F_ELOC, W * 3 * 16);
  add(A_FIELD0 + PXB_F_DETR, W * 8);
  add(A_FIELD0 + PXB_F_LOG_DETR, W * 8);
  add(A_FIELD0 + PXB_F_ESTIMATES, 10 * 16);
  add(A_FIELD0 + PXB_F_COUNTERS, 8 * 8);
  add(A_FIELD0 + PXB_F_PARENT_IX, Wt * 4);
  add(A_FIELD0 + PXB_F_XBAR, W * d.Np * 16);
  add(A_FIELD0 + PXB_F_XSHIFTED, W * d.Np * 16);
  add(A_FIELD0 + PXB_F_CMF_CFB, W * 2 * 16);
  add(A_FIELD0 + PXB_F_OVLP_NEW, W * 16);
  add(A_FIELD0 + PXB_F_TOTAL_WEIGHT, 8);
  add(A_FIELD0 + PXB_F_PAIRS, (1 + 2 * Wt) * 4);
  add(A_FIELD0 + PXB_F_PHASE, W * 16);
  h->nbp = cfg->nbp > 0 ? cfg->nbp : 0;
  const bool bp = h->nbp > 0;
  h->bp_chunks = bp ? std::max(1, std::min(64, d.W / 16)) : 0;
  add(A_FC, bp ? fc_size(d, h->nbp) * 8 : 0);
  add(A_PHI_OLD, bp ? of_size(d) * 8 : 0);
  add(A_PHI_BP, bp ? of_size(d) * 8 : 0);
  add(A_PHI_BP2, bp ? of_size(d) * 8 : 0);
  add(A_THETA_BP, bp ? of_size(d) * 8 : 0);
  add(A_BP_PART, bp ? (size_t)h->bp_chunks * 2 * d.M * d.M * 16 : 0);
  add(A_BFT, bp ? bf_size(d) * 8 : 0);
  addd(A_PSI_NAT, (size_t)d.M * d.ne * 16);
  add(A_INIT_NAT, (size_t)d.M * d.ne * 16);
  add(A_STEP_PARAMS, 256);
  addd(A_OVLP_DET, W * 16);
  addd(A_ELOC_DET, W * 3 * 16);
  add(A_XC, D > 1 ? (size_t)2 * W * d.Np * 16 : 0);
  add(A_COEFF, (size_t)PXB_MAX_DETS * 16);
  add(A_ELOC_MIX, W * 16);
  add(A_OT_TRUE, W * 16);
  add(A_BPFAC, W * 32);
  add(A_BPW, W * 16);
  add(A_FIELD0 + PXB_F_LOG_SHIFTS, 256);   // LogShifts, then the three population sums at +64 bytes
  {
    const bool cob = (d.flags & FLAG_COMPLEX_ONE_BODY) != 0;
    add(A_BF2, cob ? 2 * bf_size(d) * 8 : 0);
    add(A_PHI_STACK, (cob || kmul(d) == 2) ? 2 * of_size(d) * 8 : 0);
  }
  add(A_FIELD0 + PXB_F_WALKER_ELOC, W * 16);
  add(A_FIELD0 + PXB_F_OVLP_DET, 0);  // alias of A_OVLP_DET, fixed up below
  add(A_FIELD0 + PXB_F_BP_RDM, bp ? (size_t)2 * d.M * d.M * 16 : 0);
  add(A_FIELD0 + PXB_F_BP_DENOM, 16);
  add(A_FIELD0 + PXB_F_THETA_SUM, (size_t)d.ne * d.M * 16);
  h->reg[A_FIELD0 + PXB_F_OVLP_DET] = h->reg[A_OVLP_DET];
  h->reg[A_FIELD0 + PXB_F_OVLP_DET].bytes = h->reg[A_OVLP_DET].stride * (size_t)D;
  h->arena_bytes = off;
  *out = h;
  return PXB_OK;
}

int pxb_destroy(pxb_handle h) {
  if (h) {
    for (auto& p : h->pending) {
      cudaEventDestroy(p.a);
      cudaEventDestroy(p.b);
    }
    for (auto e : h->evpool) cudaEventDestroy(e);
    for (int r = 0; r < PXB_MAX_PEERS; ++r)
      if (h->peer_map[r]) cudaIpcCloseMemHandle(h->peer_map[r]);
    for (auto& g : h->graphs)
      if (g.exec) cudaGraphExecDestroy(g.exec);
    if (h->ev_fork) cudaEventDestroy(h->ev_fork);
    if (h->ev_join) cudaEventDestroy(h->ev_join);
    if (h->ev_in) cudaEventDestroy(h->ev_in);
    if (h->ev_out) cudaEventDestroy(h->ev_out);
    if (h->side) cudaStreamDestroy(h->side);
    if (h->main) cudaStreamDestroy(h->main);
  }
  delete h;
  return PXB_OK;
}

int pxb_arena_bytes(pxb_handle h, size_t* bytes) {
  if (!h || !bytes) return PXB_ERR_ARG;
  *bytes = h->arena_bytes;
  return PXB_OK;
}

int pxb_bind_arena(pxb_handle h, void* dev_arena, size_t bytes, void* stream) {
  if (!h || !dev_arena) return PXB_ERR_ARG;
  if (bytes < h->arena_bytes) return fail(h, PXB_ERR_ARG, "arena too small");
  if ((reinterpret_cast<uintptr_t>(dev_arena) & 255) != 0) return fail(h, PXB_ERR_ARG, "arena must be 256-byte aligned");
  PXB_CUDA(h, cudaSetDevice(h->cfg.device));
  PXB_CUDA(h, cudaMemsetAsync(dev_arena, 0, h->arena_bytes, S(stream)));
  h->arena = static_cast<unsigned char*>(dev_arena);
  if (h->d.Wtot == h->d.W) {  // one device: the "peer" table is just this arena
    h->peer_rank = 0;
    h->peer_n = 1;
    h->peer_base[0] = h->arena;
  }
  h->ham_set = false;
  h->ham_base_set = false;
  h->dets_set = 0;
  h->phi_cur = 0;
  h->theta_valid = h->x_valid = h->eloc_valid = false;
  return PXB_OK;
}

int pxb_field(pxb_handle h, int field_id, size_t* offset_bytes, size_t* size_bytes) {
  if (!h || field_id < 0 || field_id >= PXB_F_COUNT) return PXB_ERR_ARG;
  if (offset_bytes) *offset_bytes = h->reg[A_FIELD0 + field_id].off;
  if (size_bytes) *size_bytes = h->reg[A_FIELD0 + field_id].bytes;
  return PXB_OK;
}

// trial-dependent operands of determinant `det`: half-rotated Cholesky vectors (RF), psi as overlap
// GEMM fragments (PF), half-rotated one-body integrals, half-rotated ERI (K) when the exchange uses it
static int set_trial_det_impl(pxb_handle h, int det, const void* rchol, const void* h1rot, const void* psi,
                              const void* mf_shift, cudaStream_t st) {
  const Dims& d = h->d;
  int* flag = h->ptr0<int>(A_FLAG);
  PXB_CUDA(h, cudaMemsetAsync(flag, 0, 4, st));
  const int det_saved = h->det;
  h->det = det;
  struct Restore {
    pxb_handle h;
    int det;
    ~Restore() { h->det = det; }
  } restore{h, det_saved};
  ++h->launches;
  pack_rf_kernel<<<grid_for(rf_size(d)), 256, 0, st>>>(static_cast<const double2*>(rchol),
                                                       h->ptr<double>(A_RF), d, flag);
  ++h->launches;
  pack_small_kernel<<<grid_for((size_t)d.ne * d.Mp), 256, 0, st>>>(
      static_cast<const double2*>(psi), static_cast<const double2*>(h1rot),
      static_cast<const double2*>(mf_shift), h->ptr<double>(A_PSIT), h->ptr<double2>(A_H1ROT),
      h->ptr0<double2>(A_VBAR), d, flag);
  ++h->launches;
  pack_pf_kernel<<<grid_for((size_t)d.ne * d.Mp), 256, 0, st>>>(static_cast<const double2*>(psi),
                                                                h->ptr<double>(A_PF), d);
  PXB_CUDA(h, cudaGetLastError());
  if (h->exx_eri) {
    // identical spin blocks of R (RHF-type determinant): one K serves both spins
    const bool cmp = d.na == d.nb && d.nb > 0;
    if (cmp) {
      ++h->launches;
      rf_spin_compare_kernel<<<grid_for(rf_spin_base(d, 1)), 256, 0, st>>>(h->ptr<double>(A_RF), rf_spin_base(d, 1),
                                                                         rf_spin_base(d, 1), flag);
    }
    int f2 = 0;
    PXB_CUDA(h, cudaMemcpyAsync(&f2, flag, 4, cudaMemcpyDeviceToHost, st));
    PXB_CUDA(h, cudaStreamSynchronize(st));
    h->kf_shared[det] = cmp && (f2 & 8) == 0;
    const bool cc = kmul(d) == 2;
    for (int s = 0; s < (h->kf_shared[det] ? 1 : 2); ++s) {
      const int ns = s ? d.nb : d.na;
      if (ns == 0) continue;
      for (int part = cc ? 1 : 0; part <= (cc ? 2 : 0); ++part) {
        double* KF = part == 2 ? h->ptr<double>(s ? A_KF1I : A_KF0I) : h->ptr<double>(s ? A_KF1 : A_KF0);
        PXB_CUDA(h, cudaMemsetAsync(KF, 0, eri_kf_doubles(d, s) * 8, st));
        const int tiles = (d.M + 31) / 32;
        ++h->launches;
        eri_build_kernel<<<dim3(tiles * tiles, ns, ns), 1024, 0, st>>>(static_cast<const double2*>(rchol), KF, d, s,
                                                                       part);
        PXB_CUDA(h, cudaGetLastError());
      }
    }
  }
  int hflag = 0;
  PXB_CUDA(h, cudaMemcpyAsync(&hflag, flag, 4, cudaMemcpyDeviceToHost, st));
  PXB_CUDA(h, cudaStreamSynchronize(st));
  hflag &= 5;  // bit 0: rchol, bit 2: psi (bit 1 is bh1, bit 3 the spin-block comparison of the ERI setup)
  if (kmul(d) == 2) hflag = 0;  // PXB_FLAG_COMPLEX_CHOLESKY: both may be complex
  if (hflag != 0) {
    char buf[160];
    snprintf(buf, sizeof buf,
             "complex-valued Cholesky vectors / trial orbitals: create the handle with PXB_FLAG_COMPLEX_CHOLESKY "
             "(determinant %d: rchol:%d psi:%d)",
             det, hflag & 1, (hflag >> 2) & 1);
    return fail(h, PXB_ERR_UNSUPPORTED, buf);
  }
  PXB_CUDA(h, cudaMemcpyAsync(h->ptr<void>(A_PSI_NAT), psi, (size_t)d.M * d.ne * 16, cudaMemcpyDeviceToDevice, st));
  h->dets_set |= 1u << det;
  h->theta_valid = h->x_valid = h->eloc_valid = false;
  return PXB_OK;
}

int pxb_set_hamiltonian(pxb_handle h, const double* hs_pot, const void* rchol, const void* bh1,
                        const void* h1rot, const void* psi, const void* mf_shift, double ecore,
                        void* stream) {
  if (!h) return PXB_ERR_ARG;
  if (!h->arena) return fail(h, PXB_ERR_STATE, "arena not bound");
  if (!hs_pot || !rchol || !bh1 || !h1rot || !psi || !mf_shift) return fail(h, PXB_ERR_ARG, "null pointer");
  const Dims& d = h->d;
  cudaStream_t st = S(stream);
  int* flag = h->ptr0<int>(A_FLAG);
  PXB_CUDA(h, cudaMemsetAsync(flag, 0, 4, st));
  ++h->launches;
  hs_symmetry_kernel<<<grid_for((size_t)d.M * d.M * d.N), 256, 0, st>>>(hs_pot, d, flag);
  ++h->launches;
  pack_bf_kernel<<<grid_for(bf_size(d)), 256, 0, st>>>(static_cast<const double2*>(bh1), h->ptr<double>(A_BF),
                                                       d, flag, 0);
  if (h->nbp > 0) {  // BH1^T for the B^dagger chain of the back propagation
    ++h->launches;
    pack_bf_kernel<<<grid_for(bf_size(d)), 256, 0, st>>>(static_cast<const double2*>(bh1), h->ptr<double>(A_BFT),
                                                         d, flag, 1);
  }
  PXB_CUDA(h, cudaGetLastError());
  {
    // symmetric Cholesky matrices (real orbitals): keep the row tiles (p pair, q chunk) that touch
    // the upper triangle, 4 kc + 3 >= 2 pp; the GEMM epilogue mirrors them
    int f1 = 0;
    PXB_CUDA(h, cudaMemcpyAsync(&f1, flag, 4, cudaMemcpyDeviceToHost, st));
    PXB_CUDA(h, cudaStreamSynchronize(st));
    if ((f1 & 2) && !(d.flags & FLAG_COMPLEX_ONE_BODY))
      return fail(h, PXB_ERR_UNSUPPORTED,
                  "complex-valued one-body propagator: create the handle with PXB_FLAG_COMPLEX_ONE_BODY");
    if (d.flags & FLAG_COMPLEX_ONE_BODY) {
      ++h->launches;
      pack_bf2_kernel<<<grid_for(2 * bf_size(d)), 256, 0, st>>>(static_cast<const double2*>(bh1),
                                                               h->ptr<double>(A_BF2), d);
      PXB_CUDA(h, cudaGetLastError());
    }
    h->vhs_sym = h->vhs_sym_allowed && (f1 & 16) == 0;
    h->hs_near_sym = (f1 & 32) == 0;
    const int* map = nullptr;
    int nrt = d.RT;
    if (h->vhs_sym) {
      std::vector<int> m;
      const int npp = (d.M + 1) / 2;
      for (int pp = 0; pp < npp; ++pp)
        for (int kc = 0; kc < d.KC; ++kc)
          if (4 * kc + 3 >= 2 * pp) m.push_back(pp * d.KC + kc);
      h->rtu = nrt = (int)m.size();
      PXB_CUDA(h, cudaMemcpyAsync(h->ptr<int>(A_RTMAP), m.data(), m.size() * 4, cudaMemcpyHostToDevice, st));
      PXB_CUDA(h, cudaStreamSynchronize(st));
      map = h->ptr<int>(A_RTMAP);
    }
    ++h->launches;
    pack_lf_kernel<<<grid_for((size_t)nrt * d.NKC * 32), 256, 0, st>>>(hs_pot, h->ptr<double>(A_LF), d, map, nrt);
    PXB_CUDA(h, cudaGetLastError());
  }
  h->dets_set = 0;
  int rc = set_trial_det_impl(h, 0, rchol, h1rot, psi, mf_shift, st);
  if (rc) return rc;
  {  // CI coefficients default to 1
    std::vector<double> ones(2 * PXB_MAX_DETS, 0.0);
    for (int i = 0; i < PXB_MAX_DETS; ++i) ones[2 * i] = 1.0;
    PXB_CUDA(h, cudaMemcpyAsync(h->ptr0<void>(A_COEFF), ones.data(), ones.size() * 8, cudaMemcpyHostToDevice, st));
    PXB_CUDA(h, cudaStreamSynchronize(st));
  }
  h->d.ecore = ecore;
  h->ham_set = h->dets_set == (1u << h->ndets) - 1u;
  h->ham_base_set = true;
  h->theta_valid = h->x_valid = h->eloc_valid = false;
  return PXB_OK;
}

int pxb_set_trial_det(pxb_handle h, int det, double coeff_re, double coeff_im, const void* rchol,
                      const void* h1rot, const void* psi, void* stream) {
  if (!h) return PXB_ERR_ARG;
  if (!h->arena) return fail(h, PXB_ERR_STATE, "arena not bound");
  if (!h->ham_base_set) return fail(h, PXB_ERR_STATE, "pxb_set_hamiltonian has to come first");
  if (det < 0 || det >= h->ndets) return fail(h, PXB_ERR_ARG, "pxb_set_trial_det: determinant index out of range");
  cudaStream_t st = S(stream);
  if (rchol || h1rot || psi) {
    if (!rchol || !h1rot || !psi) return fail(h, PXB_ERR_ARG, "pxb_set_trial_det: rchol, h1rot and psi go together");
    int rc = set_trial_det_impl(h, det, rchol, h1rot, psi, nullptr, st);
    if (rc) return rc;
  }
  const double c[2] = {coeff_re, coeff_im};
  PXB_CUDA(h, cudaMemcpyAsync(h->ptr0<double>(A_COEFF) + 2 * det, c, 16, cudaMemcpyHostToDevice, st));
  PXB_CUDA(h, cudaStreamSynchronize(st));
  h->ham_set = h->dets_set == (1u << h->ndets) - 1u;
  h->theta_valid = h->x_valid = h->eloc_valid = false;
  return PXB_OK;
}

int pxb_set_eshift_imag(pxb_handle h, double eshift_im) {
  if (!h) return PXB_ERR_ARG;
  h->eshift_im = eshift_im;
  return PXB_OK;
}

int pxb_set_phi(pxb_handle h, const void* dev_phi, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  ++h->launches;
  phi_to_of_kernel<<<grid_for((size_t)d.Wp * d.ne * d.Mp), 256, 0, S(stream)>>>(
      static_cast<const double2*>(dev_phi), h->phi(), d, 0);
  PXB_CUDA(h, cudaGetLastError());
  h->theta_valid = h->x_valid = h->eloc_valid = false;
  return PXB_OK;
}

int pxb_get_phi(pxb_handle h, void* dev_phi, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  ++h->launches;
  of_to_natural_kernel<<<grid_for((size_t)d.W * d.ne * d.M), 256, 0, S(stream)>>>(
      h->phi(), static_cast<double2*>(dev_phi), d, 1);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_init_walkers(pxb_handle h, const void* dev_init_phi, double total_walkers, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  cudaStream_t st = S(stream);
  ++h->launches;
  phi_to_of_kernel<<<grid_for((size_t)d.Wp * d.ne * d.Mp), 256, 0, st>>>(
      static_cast<const double2*>(dev_init_phi), h->phi(), d, 1);
  PXB_CUDA(h, cudaGetLastError());
  h->theta_valid = h->x_valid = h->eloc_valid = false;
  int rc = ensure_theta(h, st);
  if (rc) return rc;
  ++h->launches;
  init_scalars_kernel<<<(d.Wp + 255) / 256, 256, 0, st>>>(
      h->field<double>(PXB_F_WEIGHT), h->field<double>(PXB_F_UNSCALED_WEIGHT), h->field<double2>(PXB_F_OT),
      h->ptr<double2>(A_OVLP_OLD), h->field<double2>(PXB_F_HYBRID_ENERGY), h->field<double>(PXB_F_DETR),
      h->field<double>(PXB_F_LOG_DETR), h->field<double>(PXB_F_TOTAL_WEIGHT),
      h->field<double2>(PXB_F_PHASE), h->field<double2>(PXB_F_WALKER_ELOC), h->ptr0<double2>(A_OT_TRUE),
      total_walkers, d);
  PXB_CUDA(h, cudaGetLastError());
  ++h->launches;
  bpfac_reset_kernel<<<(d.Wp + 255) / 256, 256, 0, st>>>(h->ptr0<double2>(A_BPFAC), d.Wp);
  PXB_CUDA(h, cudaGetLastError());
  {  // handler.py:46 shift_counter = 1, all shifts zero; `enabled` is kept
    LogShifts ls{0.0, 0.0, 0.0, 1, h->log_shift_on ? 1 : 0, 0};
    PXB_CUDA(h, cudaMemcpyAsync(h->field<void>(PXB_F_LOG_SHIFTS), &ls, sizeof ls, cudaMemcpyHostToDevice, st));
    PXB_CUDA(h, cudaStreamSynchronize(st));
  }
  PXB_CUDA(h, cudaGetLastError());
  PXB_CUDA(h, cudaMemsetAsync(h->field<void>(PXB_F_ESTIMATES), 0, 160, st));
  PXB_CUDA(h, cudaMemsetAsync(h->field<void>(PXB_F_COUNTERS), 0, 64, st));
  PXB_CUDA(h, cudaMemsetAsync(h->field<void>(PXB_F_ELOC), 0, (size_t)d.Wp * 48, st));
  PXB_CUDA(h, cudaMemcpyAsync(h->ptr<void>(A_INIT_NAT), dev_init_phi, (size_t)d.M * d.ne * 16,
                              cudaMemcpyDeviceToDevice, st));
  if (h->nbp > 0) {
    // walker.phi_old = phi.copy() (walkers/walker.py:43), empty field history
    PXB_CUDA(h, cudaMemcpyAsync(h->ptr<void>(A_PHI_OLD), h->phi(), of_size(d) * 8, cudaMemcpyDeviceToDevice, st));
    PXB_CUDA(h, cudaMemsetAsync(h->ptr<void>(A_FC), 0, fc_size(d, h->nbp) * 8, st));
    PXB_CUDA(h, cudaMemsetAsync(h->field<void>(PXB_F_BP_RDM), 0, (size_t)2 * d.M * d.M * 16, st));
    PXB_CUDA(h, cudaMemsetAsync(h->field<void>(PXB_F_BP_DENOM), 0, 16, st));
    h->bp_step = 0;
  }
  return PXB_OK;
}

int pxb_propagate(pxb_handle h, const double* dev_xi, uint64_t rng_seed, int64_t walker_offset,
                  double eshift, int64_t step, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  cudaStream_t st = S(stream);
  int* active = h->ptr<int>(A_ACTIVE);
  long long* counters = h->field<long long>(PXB_F_COUNTERS);
  int rc;
  if (!h->in_step) {
    StepParams sp{rng_seed, (unsigned long long)step, walker_offset, eshift, h->eshift_im, 0.0};
    if ((rc = set_step_params(h, sp, 1, st))) return rc;
  }
  ++h->launches;
  active_kernel<<<(d.Wp + 255) / 256, 256, 0, st>>>(h->field<double>(PXB_F_WEIGHT), active, counters, d);
  PXB_CUDA(h, cudaGetLastError());
  // (a) Theta of the current walkers (normally still valid from the previous step's
  //     closing Green's function / the estimator); ovlp_old is walker.ot
  // the reference recomputes the overlap at the top of propagate_walker_phaseless
  // (continuous.py:245); walker.ot stands in for it only while Theta is current -- after
  // pxb_set_phi / a restart read the freshly computed overlap (A_OVLP_OLD) is used instead
  const bool ot_stale = !h->theta_valid;
  if ((rc = ensure_theta(h, st))) return rc;
  // (c1) force bias GEMM X_s = R_s^T Theta_s (shared with the Coulomb term of the estimator)
  if ((rc = ensure_x(h, st))) return rc;
  const bool le_mode = (d.flags & FLAG_LOCAL_ENERGY_WEIGHT) != 0;
  // local-energy weight update: walker.local_energy of continuous.py:296 uses the Green's functions
  // of the walker BEFORE the step (left by greens_function at the top of it)
  if (le_mode && (rc = run_det_energies(h, false, st))) return rc;
  FieldArgs f;
  f.X = h->ndets > 1 ? h->ptr0<double2>(A_XC) : h->ptr0<double2>(A_X);
  f.xi = dev_xi;
  f.vbar = h->ptr<double2>(A_VBAR);
  f.active = active;
  f.XF = h->ptr<double>(A_XF);
  f.xbar_out = h->field<double2>(PXB_F_XBAR);
  f.xs_out = h->field<double2>(PXB_F_XSHIFTED);
  f.cmfcfb = h->field<double2>(PXB_F_CMF_CFB);
  f.counters = counters;
  f.d = d;
  f.sp = h->ptr<StepParams>(A_STEP_PARAMS);
  {
    StageTimer timer__(h, PXB_STAGE_FIELD, st);
    ++h->launches;
    field_kernel<<<(d.Wp + 7) / 8, 256, 0, st>>>(f);
  }
  PXB_CUDA(h, cudaGetLastError());
  if (h->nbp > 0 && !(d.flags & FLAG_FREE_PROJECTION)) {
    // FieldConfig.update (walkers/stack.py:52-79, called from continuous.py:288-289): keep x
    if (h->bp_step >= h->nbp)
      return fail(h, PXB_ERR_STATE, "field history is full: call pxb_bp_reset after back propagating");
    const size_t row = (size_t)d.NKC * 32 * 8;
    PXB_CUDA(h, cudaMemcpy2DAsync(h->ptr<unsigned char>(A_FC) + (size_t)h->bp_step * row, row * h->nbp,
                                  h->ptr<void>(A_XF), row, row, d.WG, cudaMemcpyDeviceToDevice, st));
    ++h->bp_step;
  }
  if ((rc = run_vhs_gemm(h, st))) return rc;
  double* work = h->phi_other();
  if ((rc = run_one_body(h, h->phi(), work, nullptr, st))) return rc;
  if ((rc = run_taylor(h, work, active, st))) return rc;
  // (d) second half step writes back into the walker buffer, active walkers only
  if ((rc = run_one_body(h, work, h->phi(), active, st))) return rc;
  // (e) Green's function of the propagated walkers: its determinant is the new overlap
  //     (single_det.py:170-199) and its Theta serves the estimator and the next step
  h->theta_valid = h->x_valid = h->eloc_valid = false;
  if ((rc = run_greens_all(h, h->phi(), h->field<double2>(PXB_F_OVLP_NEW), st))) return rc;
  h->theta_valid = true;
  // (f) weights
  if (le_mode) {
    // eloc = sum_i w_i(new) E_i(old) / sum_i w_i(new): calc_overlap refreshed the determinant
    // weights, the Green's functions are still those of the start of the step (continuous.py:296)
    if ((rc = run_md_energy(h, h->ptr0<double2>(A_ELOC_MIX), 1, st))) return rc;
    WeightLeArgs wl;
    wl.weight = h->field<double>(PXB_F_WEIGHT);
    wl.ot = h->field<double2>(PXB_F_OT);
    wl.walker_eloc = h->field<double2>(PXB_F_WALKER_ELOC);
    wl.ot_true = h->ptr0<double2>(A_OT_TRUE);
    wl.shifts = h->field<LogShifts>(PXB_F_LOG_SHIFTS);
    wl.eloc_mix = h->ptr0<double2>(A_ELOC_MIX);
    wl.ovlp_new = h->field<double2>(PXB_F_OVLP_NEW);
    wl.ovlp_old = ot_stale ? h->ptr0<double2>(A_OVLP_OLD)
                           : (h->log_shift_on ? h->ptr0<double2>(A_OT_TRUE) : h->field<double2>(PXB_F_OT));
    wl.active = active;
    wl.total_weight = h->field<double>(PXB_F_TOTAL_WEIGHT);
    wl.counters = counters;
    wl.d = d;
    wl.sp = h->ptr0<StepParams>(A_STEP_PARAMS);
    StageTimer timer__(h, PXB_STAGE_WEIGHT, st);
    ++h->launches;
    weight_le_kernel<<<(d.W + 255) / 256, 256, 0, st>>>(wl);
    PXB_CUDA(h, cudaGetLastError());
    return PXB_OK;
  }
  WeightArgs wa;
  wa.weight = h->field<double>(PXB_F_WEIGHT);
  wa.phase = h->field<double2>(PXB_F_PHASE);
  wa.ot = h->field<double2>(PXB_F_OT);
  wa.ehyb = h->field<double2>(PXB_F_HYBRID_ENERGY);
  wa.ovlp_new = h->field<double2>(PXB_F_OVLP_NEW);
  wa.ovlp_old = ot_stale ? h->ptr0<double2>(A_OVLP_OLD)
                         : (h->log_shift_on ? h->ptr0<double2>(A_OT_TRUE) : h->field<double2>(PXB_F_OT));
  wa.ot_true = h->ptr0<double2>(A_OT_TRUE);
  wa.shifts = h->field<LogShifts>(PXB_F_LOG_SHIFTS);
  wa.bpfac = (h->nbp > 0 && !(d.flags & FLAG_FREE_PROJECTION)) ? h->ptr0<double2>(A_BPFAC) : nullptr;
  wa.cmfcfb = h->field<double2>(PXB_F_CMF_CFB);
  wa.active = active;
  wa.total_weight = h->field<double>(PXB_F_TOTAL_WEIGHT);
  wa.counters = counters;
  wa.d = d;
  wa.sp = h->ptr0<StepParams>(A_STEP_PARAMS);
  {
    StageTimer timer__(h, PXB_STAGE_WEIGHT, st);
    ++h->launches;
    weight_kernel<<<(d.W + 255) / 256, 256, 0, st>>>(wa);
  }
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_orthogonalise(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  cudaStream_t st = S(stream);
  StageTimer timer__(h, PXB_STAGE_QR, st);
  {
    const int rc = run_qr(h, h->phi(), st);
    if (rc) return rc;
  }
  double* qrld = h->ptr<double>(A_QRLD);
  ++h->launches;
  qr_combine_kernel<<<(d.W + 255) / 256, 256, 0, st>>>(
      qrld, h->field<double2>(PXB_F_OT), h->ptr0<double2>(A_OT_TRUE), h->field<double>(PXB_F_DETR),
      h->field<double>(PXB_F_LOG_DETR), (d.flags & FLAG_FREE_PROJECTION) ? h->field<double>(PXB_F_WEIGHT) : nullptr,
      h->log_shift_on ? &h->field<LogShifts>(PXB_F_LOG_SHIFTS)->detR_shift : nullptr, d.W);
  PXB_CUDA(h, cudaGetLastError());
  // Theta = O^-1 phi^T is invariant under phi -> phi R^-1: it stays valid (to rounding)
  return PXB_OK;
}

int pxb_local_energy(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  cudaStream_t st = S(stream);
  int rc;
  // already evaluated for these walkers (e.g. before the population control: ELOC travels with
  // the walker payload, so the estimator's call after it finds them in place)
  if (h->theta_valid && h->x_valid && h->eloc_valid) return PXB_OK;
  if ((rc = ensure_theta(h, st))) return rc;
  if ((rc = ensure_x(h, st))) return rc;
  if ((rc = run_det_energies(h, true, st))) return rc;
  if (h->ndets > 1 && (rc = run_md_energy(h, h->field<double2>(PXB_F_ELOC), 0, st))) return rc;
  h->eloc_valid = true;
  return PXB_OK;
}

int pxb_accumulate(pxb_handle h, int with_energy, void* stream) {
  PXB_REQUIRE_READY(h);
  AccArgs a;
  a.weight = h->field<double>(PXB_F_WEIGHT);
  a.phase = h->field<double2>(PXB_F_PHASE);
  a.unscaled = h->field<double>(PXB_F_UNSCALED_WEIGHT);
  a.ot = h->field<double2>(PXB_F_OT);
  a.ehyb = h->field<double2>(PXB_F_HYBRID_ENERGY);
  a.eloc = h->field<double2>(PXB_F_ELOC);
  a.estimates = h->field<double2>(PXB_F_ESTIMATES);
  a.d = h->d;
  a.with_energy = with_energy;
  StageTimer timer__(h, PXB_STAGE_ACCUMULATE, S(stream));
  ++h->launches;
  if (h->d.flags & FLAG_FREE_PROJECTION)
    accumulate_free_kernel<<<1, 1024, 0, S(stream)>>>(a);
  else
    accumulate_kernel<<<1, 1024, 0, S(stream)>>>(a);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_accumulate_theta(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  if (h->ndets > 1) return fail(h, PXB_ERR_UNSUPPORTED, "mixed one-body density matrix: single-determinant trials only");
  int rc = ensure_theta(h, S(stream));
  if (rc) return rc;
  StageTimer timer__(h, PXB_STAGE_ACCUMULATE, S(stream));
  ++h->launches;
  theta_wsum_kernel<<<d.ne * d.KC, 128, 0, S(stream)>>>(h->ptr<double>(A_THETA), h->field<double>(PXB_F_WEIGHT),
                                                       h->field<double2>(PXB_F_THETA_SUM), d);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_zero_estimates(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  PXB_CUDA(h, cudaMemsetAsync(h->field<void>(PXB_F_THETA_SUM), 0, (size_t)h->d.ne * h->d.M * 16, S(stream)));
  PXB_CUDA(h, cudaMemsetAsync(h->field<void>(PXB_F_ESTIMATES), 0, 160, S(stream)));
  return PXB_OK;
}

// walkers.use_log_shift (walkers/handler.py:228,456-475)
int pxb_log_shift_enable(pxb_handle h, int enable, void* stream) {
  PXB_REQUIRE_READY(h);
  h->log_shift_on = enable != 0;
  const int e = enable ? 1 : 0;
  PXB_CUDA(h, cudaMemcpyAsync(&h->field<LogShifts>(PXB_F_LOG_SHIFTS)->enabled, &e, 4, cudaMemcpyHostToDevice, S(stream)));
  PXB_CUDA(h, cudaStreamSynchronize(S(stream)));
  return PXB_OK;
}

int pxb_log_shift_sums(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  ++h->launches;
  shift_sums_kernel<<<1, 1024, 0, S(stream)>>>(h->field<double2>(PXB_F_OT), h->field<double>(PXB_F_DETR),
                                               h->field<double>(PXB_F_LOG_DETR),
                                               h->field<double>(PXB_F_LOG_SHIFTS) + 8, h->d.W);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_log_shift_update(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  ++h->launches;
  shift_update_kernel<<<1, 1, 0, S(stream)>>>(h->field<LogShifts>(PXB_F_LOG_SHIFTS),
                                             h->field<double>(PXB_F_LOG_SHIFTS) + 8, (double)h->d.Wtot);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

static int pop_rescale_impl(pxb_handle h, const double* gw, int64_t wtot, int write_local, void* stream) {
  PXB_REQUIRE_READY(h);
  if (wtot != h->d.Wtot) return fail(h, PXB_ERR_ARG, "wtot != total_walkers given at create");
  RescaleArgs a;
  a.write_local = write_local;
  a.gw = gw;
  a.gws = h->ptr<double>(A_GWS);
  a.weight = h->field<double>(PXB_F_WEIGHT);
  a.unscaled = h->field<double>(PXB_F_UNSCALED_WEIGHT);
  a.total_weight = h->field<double>(PXB_F_TOTAL_WEIGHT);
  a.counters = h->field<long long>(PXB_F_COUNTERS);
  a.W = h->d.W;
  a.Wtot = (int)wtot;
  ++h->launches;
  pop_rescale_kernel<<<1, 1024, 0, S(stream)>>>(a);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_pop_rescale(pxb_handle h, const double* gw, int64_t wtot, void* stream) {
  return pop_rescale_impl(h, gw, wtot, 1, stream);
}

int pxb_comb_plan(pxb_handle h, const double* gw, int64_t wtot, double r, void* stream) {
  PXB_REQUIRE_READY(h);
  (void)gw;  // the rescaled copy written by pxb_pop_rescale is what the comb runs on
  if (wtot != h->d.Wtot) return fail(h, PXB_ERR_ARG, "wtot != total_walkers given at create");
  CombArgs a;
  a.gws = h->ptr<double>(A_GWS);
  a.cprobs = h->ptr<double>(A_CPROBS);
  a.parent_ix = h->field<int>(PXB_F_PARENT_IX);
  a.pairs = h->field<int>(PXB_F_PAIRS);
  a.counters = h->field<long long>(PXB_F_COUNTERS);
  a.Wtot = (int)wtot;
  a.sp = h->ptr<StepParams>(A_STEP_PARAMS);
  if (!h->in_step) {
    StepParams sp{0, 0, 0, 0.0, 0.0, r};
    int rc = set_step_params(h, sp, 2, S(stream));
    if (rc) return rc;
  }
  ++h->launches;
  comb_plan_kernel<<<1, 1024, 0, S(stream)>>>(a);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_set_weights(pxb_handle h, double value, void* stream) {
  PXB_REQUIRE_READY(h);
  ++h->launches;
  fill_kernel<<<(h->d.W + 255) / 256, 256, 0, S(stream)>>>(h->field<double>(PXB_F_WEIGHT), value, h->d.W,
                                                           h->field<long long>(PXB_F_COUNTERS) + 4);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_pop_control_comb(pxb_handle h, double r, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  if (d.Wtot != d.W) return fail(h, PXB_ERR_ARG, "pxb_pop_control_comb is the single-device path");
  if (d.W == 1) return PXB_OK;  // handler.py:226-227
  cudaStream_t st = S(stream);
  StageTimer timer__(h, PXB_STAGE_POP_CONTROL, st);
  double* gw = h->ptr<double>(A_GW);
  ++h->launches;
  abs_weight_kernel<<<(d.W + 255) / 256, 256, 0, st>>>(h->field<double>(PXB_F_WEIGHT), gw, d.W);
  PXB_CUDA(h, cudaGetLastError());
  int rc;
  if ((rc = pxb_pop_rescale(h, gw, d.W, stream))) return rc;
  if ((rc = pxb_comb_plan(h, gw, d.W, r, stream))) return rc;
  ++h->launches;
  copy_pairs_kernel<<<std::min(d.W, 4 * h->sm_count), 256, 0, st>>>(copy_args(h), h->field<int>(PXB_F_PAIRS), 0);
  PXB_CUDA(h, cudaGetLastError());
  if (h->ndets > 1) h->theta_valid = h->x_valid = false;  // only determinant 0's Theta / X travel
  return pxb_set_weights(h, 1.0, stream);
}

// ---- peer-memory comb (multi-device, no host round trip) ----------------------
int pxb_peer_export(pxb_handle h, void* handle_out, uint64_t* offset_out) {
  if (!h || !handle_out || !offset_out) return PXB_ERR_ARG;
  if (!h->arena) return fail(h, PXB_ERR_STATE, "arena not bound");
  static_assert(sizeof(cudaIpcMemHandle_t) == PXB_IPC_HANDLE_BYTES, "IPC handle size");
  PXB_CUDA(h, cudaSetDevice(h->cfg.device));
  // the arena may sit inside a larger cudaMalloc block (caching allocators): the IPC handle
  // names the block, so the offset of the arena inside it travels with the handle
  typedef int (*get_range_fn)(unsigned long long*, size_t*, unsigned long long);
  void* fp = nullptr;
  cudaDriverEntryPointQueryResult qr;
  PXB_CUDA(h, cudaGetDriverEntryPoint("cuMemGetAddressRange", &fp, cudaEnableDefault, &qr));
  if (!fp || qr != cudaDriverEntryPointSuccess) return fail(h, PXB_ERR_CUDA, "cuMemGetAddressRange not found");
  unsigned long long base = 0;
  size_t size = 0;
  if (reinterpret_cast<get_range_fn>(fp)(&base, &size, (unsigned long long)(uintptr_t)h->arena) != 0)
    return fail(h, PXB_ERR_CUDA, "cuMemGetAddressRange failed");
  cudaIpcMemHandle_t mh;
  PXB_CUDA(h, cudaIpcGetMemHandle(&mh, reinterpret_cast<void*>((uintptr_t)base)));
  std::memcpy(handle_out, &mh, sizeof(mh));
  *offset_out = (uint64_t)((uintptr_t)h->arena - (uintptr_t)base);
  return PXB_OK;
}

int pxb_peer_attach(pxb_handle h, int rank, int nranks, const void* handles, const uint64_t* offsets) {
  if (!h || !handles || !offsets) return PXB_ERR_ARG;
  if (!h->arena) return fail(h, PXB_ERR_STATE, "arena not bound");
  if (nranks < 1 || nranks > PXB_MAX_PEERS || rank < 0 || rank >= nranks)
    return fail(h, PXB_ERR_ARG, "pxb_peer_attach: bad rank / nranks");
  if ((long long)nranks * h->d.W != h->d.Wtot)
    return fail(h, PXB_ERR_ARG, "pxb_peer_attach: nranks * nwalkers != total_walkers");
  PXB_CUDA(h, cudaSetDevice(h->cfg.device));
  const unsigned char* hb = static_cast<const unsigned char*>(handles);
  for (int r = 0; r < nranks; ++r) {
    if (r == rank) {
      h->peer_base[r] = h->arena;
      continue;
    }
    cudaIpcMemHandle_t mh;
    std::memcpy(&mh, hb + (size_t)r * PXB_IPC_HANDLE_BYTES, sizeof(mh));
    void* p = nullptr;
    PXB_CUDA(h, cudaIpcOpenMemHandle(&p, mh, cudaIpcMemLazyEnablePeerAccess));
    h->peer_map[r] = p;
    h->peer_base[r] = static_cast<unsigned char*>(p) + offsets[r];
  }
  h->peer_rank = rank;
  h->peer_n = nranks;
  return PXB_OK;
}

// total weight + comb plan only (no walker state is written): may run on a side stream while
// pxb_local_energy works on the launch stream; gw == NULL (one device): |weight| is taken locally
int pxb_pop_plan(pxb_handle h, const double* gw, int64_t wtot, double r, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  cudaStream_t st = S(stream);
  StageTimer timer__(h, PXB_STAGE_POP_CONTROL, st);
  if (gw == nullptr) {
    if (d.Wtot != d.W) return fail(h, PXB_ERR_ARG, "pxb_pop_plan: global weights needed with several devices");
    double* lgw = h->ptr<double>(A_GW);
    ++h->launches;
    abs_weight_kernel<<<(d.W + 255) / 256, 256, 0, st>>>(h->field<double>(PXB_F_WEIGHT), lgw, d.W);
    PXB_CUDA(h, cudaGetLastError());
    gw = lgw;
    wtot = d.W;
  }
  int rc;
  if ((rc = pop_rescale_impl(h, gw, wtot, 0, stream))) return rc;
  return pxb_comb_plan(h, gw, wtot, r, stream);
}

// data movement of the plan: every killed slot of this device receives its clone (local copy or
// NVLink pull from the owner's arena)
int pxb_pop_pull(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  if (h->peer_n < 1) return fail(h, PXB_ERR_STATE, "pxb_peer_attach has not been called");
  cudaStream_t st = S(stream);
  StageTimer timer__(h, PXB_STAGE_POP_CONTROL, st);
  PeerArgs p;
  for (int i = 0; i < PXB_MAX_PEERS; ++i) p.base[i] = h->peer_base[i];
  p.rank = h->peer_rank;
  p.nranks = h->peer_n;
  p.nw = d.W;
  ++h->launches;
  pull_pairs_kernel<<<4 * h->sm_count, 256, 0, st>>>(copy_args(h), p, h->field<int>(PXB_F_PAIRS));
  PXB_CUDA(h, cudaGetLastError());
  if (h->ndets > 1) h->theta_valid = h->x_valid = false;  // only determinant 0's Theta / X travel
  return PXB_OK;  // Theta, e1b, X and ELOC travel with the walkers: their validity is unchanged
}

int pxb_pop_control_comb_peers(pxb_handle h, const double* gw, int64_t wtot, double r, void* stream) {
  int rc = pxb_pop_plan(h, gw, wtot, r, stream);
  if (rc) return rc;
  return pxb_pop_pull(h, stream);
}

int pxb_reserve_sms(pxb_handle h, int n) {
  if (!h || n < 0) return PXB_ERR_ARG;
  h->reserved_sms = n;
  return PXB_OK;
}

int pxb_pop_control_finish(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  StageTimer timer__(h, PXB_STAGE_POP_CONTROL, S(stream));
  ++h->launches;
  pop_finish_kernel<<<(d.W + 255) / 256, 256, 0, S(stream)>>>(
      h->field<double>(PXB_F_WEIGHT), h->field<double>(PXB_F_UNSCALED_WEIGHT), 1.0, d.W,
      h->field<long long>(PXB_F_COUNTERS));
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_payload_doubles(pxb_handle h, size_t* n) {
  if (!h || !n) return PXB_ERR_ARG;
  *n = payload_doubles(h->d, h->nbp > 0, (int)fc_rows(h->d, h->nbp));
  return PXB_OK;
}

int pxb_copy_walkers(pxb_handle h, const int32_t* src, const int32_t* dst, int n, void* stream) {
  PXB_REQUIRE_READY(h);
  if (n <= 0) return PXB_OK;
  ++h->launches;
  copy_list_kernel<<<std::min(n, 4 * h->sm_count), 256, 0, S(stream)>>>(copy_args(h), src, dst, n);
  PXB_CUDA(h, cudaGetLastError());
  if (h->ndets > 1) h->theta_valid = h->x_valid = false;
  return PXB_OK;
}

int pxb_pack_walkers(pxb_handle h, const int32_t* slots, int n, double* buf, void* stream) {
  PXB_REQUIRE_READY(h);
  if (n <= 0) return PXB_OK;
  ++h->launches;
  pack_kernel<<<std::min(n, 4 * h->sm_count), 256, 0, S(stream)>>>(copy_args(h), slots, n, buf, 0);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_unpack_walkers(pxb_handle h, const int32_t* slots, int n, const double* buf, void* stream) {
  PXB_REQUIRE_READY(h);
  if (n <= 0) return PXB_OK;
  ++h->launches;
  pack_kernel<<<std::min(n, 4 * h->sm_count), 256, 0, S(stream)>>>(copy_args(h), slots, n,
                                                                  const_cast<double*>(buf), 1);
  PXB_CUDA(h, cudaGetLastError());
  if (h->ndets > 1) h->theta_valid = h->x_valid = false;
  return PXB_OK;
}

// handler.py:271-286 on the host, bit-exact (sequential sum and cumsum, two-pointer sweep)
int pxb_comb_plan_host(const double* weights, int64_t n, double r, int32_t* parent_ix) {
  if (!weights || !parent_ix || n < 1) return PXB_ERR_ARG;
  std::vector<double> cprobs((size_t)n);
  volatile double total = 0.0;
  for (int64_t i = 0; i < n; ++i) {
    total = total + weights[i];
    cprobs[(size_t)i] = total;
    parent_ix[i] = 0;
  }
  volatile double spacing = total / (double)n;
  int64_t iw = 0, ic = 0;
  while (ic < n) {
    volatile double s = (double)ic + r;
    volatile double tooth = s * spacing;
    if (tooth < cprobs[(size_t)iw]) {
      parent_ix[iw] += 1;
      ++ic;
    } else {
      ++iw;
      if (iw >= n) return PXB_ERR_ARG;  // the reference raises IndexError here
    }
  }
  return PXB_OK;
}

// ---- fused driver step ---------------------------------------------------------
// One pass of the loop body of qmc/afqmc.py:223-255 on ONE device as a single call:
//   [pxb_orthogonalise]  pxb_propagate  [comb: plan on a side stream || pxb_local_energy, then the
//   copies and the weight reset]  pxb_accumulate
// The sequence only depends on `flags` and on the field pointer, so it is captured as a CUDA graph
// the second time a variant is seen and replayed from then on (c1/c2-sized problems are bound by
// launch latency: ~24 launches of a few microseconds each).
namespace {
int step_sequence(pxb_handle h, const double* dev_xi, int flags, cudaStream_t st) {
  int rc;
  void* vst = reinterpret_cast<void*>(st);
  if ((flags & PXB_STEP_ORTHO) && (rc = pxb_orthogonalise(h, vst))) return rc;
  if ((rc = pxb_propagate(h, dev_xi, 0, 0, 0.0, 0, vst))) return rc;
  const bool energy = (flags & PXB_STEP_ENERGY) != 0;
  if (flags & PXB_STEP_POP) {
    if (energy) {
      // the plan needs only the weights: it runs beside the energy kernels
      PXB_CUDA(h, cudaEventRecord(h->ev_fork, st));
      PXB_CUDA(h, cudaStreamWaitEvent(h->side, h->ev_fork, 0));
      if ((rc = pxb_pop_plan(h, nullptr, 0, 0.0, reinterpret_cast<void*>(h->side)))) return rc;
      if ((rc = pxb_local_energy(h, vst))) return rc;
      PXB_CUDA(h, cudaEventRecord(h->ev_join, h->side));
      PXB_CUDA(h, cudaStreamWaitEvent(st, h->ev_join, 0));
    } else {
      if ((rc = pxb_pop_plan(h, nullptr, 0, 0.0, vst))) return rc;
    }
    if ((rc = pxb_pop_pull(h, vst))) return rc;
    if ((rc = pxb_pop_control_finish(h, vst))) return rc;
  } else if (energy) {
    if ((rc = pxb_local_energy(h, vst))) return rc;
  }
  return pxb_accumulate(h, energy ? 1 : 0, vst);
}
}  // namespace

int pxb_step(pxb_handle h, const double* dev_xi, uint64_t rng_seed, int64_t walker_offset, double eshift,
             int64_t step, double comb_r, int flags, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  if (d.Wtot != d.W) return fail(h, PXB_ERR_ARG, "pxb_step is the single-device path");
  if (h->nbp > 0) return fail(h, PXB_ERR_UNSUPPORTED, "pxb_step: not with back propagation (field history grows)");
  if ((flags & PXB_STEP_POP) && d.W == 1) flags &= ~PXB_STEP_POP;  // handler.py:226-227
  cudaStream_t st = S(stream);
  if (!h->side) {
    PXB_CUDA(h, cudaStreamCreateWithFlags(&h->side, cudaStreamNonBlocking));
    PXB_CUDA(h, cudaEventCreateWithFlags(&h->ev_fork, cudaEventDisableTiming));
    PXB_CUDA(h, cudaEventCreateWithFlags(&h->ev_join, cudaEventDisableTiming));
  }
  int rc;
  StepParams sp{rng_seed, (unsigned long long)step, walker_offset, eshift, h->eshift_im, comb_r};
  if ((rc = set_step_params(h, sp, 3, st))) return rc;
  // a replay is only valid from the state the variant was captured in
  const unsigned long long key = (unsigned long long)(flags & 0xff) | (h->theta_valid ? 0x100ull : 0) |
                                 (h->x_valid ? 0x200ull : 0) | (h->eloc_valid ? 0x400ull : 0) |
                                 ((unsigned long long)h->reserved_sms << 16);
  pxb_context::StepGraph* g = nullptr;
  if (h->graphs_enabled && !h->prof) {
    for (auto& c : h->graphs)
      if (c.key == key && c.xi == dev_xi) g = &c;
    if (!g) {
      if (h->graphs.size() >= 16) h->graphs_enabled = false;  // something varies every step: stop caching
      else {
        h->graphs.push_back({key, dev_xi, 0, nullptr, 0, false, false, false});
        g = &h->graphs.back();
      }
    }
  }
  // The legacy default stream (and the per-thread one) cannot be captured: the graph then lives on
  // an internal stream that is ordered after / before the caller's stream with two events.
  const bool own = (st == nullptr || st == cudaStreamLegacy || st == cudaStreamPerThread);
  cudaStream_t gs = st;
  if (g && own) {
    if (!h->main) {
      PXB_CUDA(h, cudaStreamCreateWithFlags(&h->main, cudaStreamNonBlocking));
      PXB_CUDA(h, cudaEventCreateWithFlags(&h->ev_in, cudaEventDisableTiming));
      PXB_CUDA(h, cudaEventCreateWithFlags(&h->ev_out, cudaEventDisableTiming));
    }
    gs = h->main;
  }
  auto enter = [&]() -> cudaError_t {
    if (gs == st) return cudaSuccess;
    cudaError_t e = cudaEventRecord(h->ev_in, st);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(gs, h->ev_in, 0);
  };
  auto leave = [&]() -> cudaError_t {
    if (gs == st) return cudaSuccess;
    cudaError_t e = cudaEventRecord(h->ev_out, gs);
    return e != cudaSuccess ? e : cudaStreamWaitEvent(st, h->ev_out, 0);
  };
  if (g && g->exec) {
    PXB_CUDA(h, enter());
    PXB_CUDA(h, cudaGraphLaunch(g->exec, gs));
    PXB_CUDA(h, leave());
    h->launches += g->kernels;
    ++h->graph_replays;
    h->theta_valid = g->theta_valid;
    h->x_valid = g->x_valid;
    h->eloc_valid = g->eloc_valid;
    return PXB_OK;
  }
  if (g && g->seen >= 1) {
    const long long l0 = h->launches;
    cudaGraph_t graph = nullptr;
    PXB_CUDA(h, enter());
    if (cudaStreamBeginCapture(gs, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      h->in_step = true;
      rc = step_sequence(h, dev_xi, flags, gs);
      h->in_step = false;
      cudaError_t ce = cudaStreamEndCapture(gs, &graph);
      if (rc == PXB_OK && ce == cudaSuccess && graph != nullptr &&
          cudaGraphInstantiate(&g->exec, graph, 0) == cudaSuccess) {
        g->kernels = h->launches - l0;
        g->theta_valid = h->theta_valid;
        g->x_valid = h->x_valid;
        g->eloc_valid = h->eloc_valid;
        cudaGraphDestroy(graph);
        PXB_CUDA(h, cudaGraphLaunch(g->exec, gs));
        PXB_CUDA(h, leave());
        ++h->graph_replays;
        return PXB_OK;
      }
      if (graph) cudaGraphDestroy(graph);
      if (rc != PXB_OK && rc != PXB_ERR_CUDA) return rc;
    }
    // capture unavailable or failed: nothing was executed; no graphs for this handle any more, the
    // step runs as plain launches from the state it was entered in
    cudaGetLastError();
    g->exec = nullptr;
    h->graphs_enabled = false;
    h->launches = l0;
    h->theta_valid = (key & 0x100ull) != 0;
    h->x_valid = (key & 0x200ull) != 0;
    h->eloc_valid = (key & 0x400ull) != 0;
  }
  h->in_step = true;
  rc = step_sequence(h, dev_xi, flags, st);
  h->in_step = false;
  if (g) ++g->seen;
  return rc;
}

int pxb_step_graphs(pxb_handle h, int enable, long long* replays) {
  if (!h) return PXB_ERR_ARG;
  if (enable >= 0) h->graphs_enabled = enable != 0;
  if (replays) *replays = h->graph_replays;
  return PXB_OK;
}

// ---- back propagation -----------------------------------------
int pxb_bp_steps(pxb_handle h) { return h ? h->bp_step : 0; }

int pxb_back_propagate(pxb_handle h, int nsteps, int nstblz, int init_walker, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  cudaStream_t st = S(stream);
  if (h->nbp <= 0) return fail(h, PXB_ERR_STATE, "back propagation was not enabled (pxb_config.nbp == 0)");
  if (nsteps < 0 || nsteps > h->bp_step) return fail(h, PXB_ERR_ARG, "pxb_back_propagate: nsteps exceeds the stored history");
  if (nstblz < 1) return fail(h, PXB_ERR_ARG, "pxb_back_propagate: nstblz < 1");
  if (!h->hs_near_sym)
    return fail(h, PXB_ERR_UNSUPPORTED, "back propagation needs symmetric Cholesky matrices (real orbitals)");
  double* phi_bp = h->ptr<double>(A_PHI_BP);
  double* work = h->ptr<double>(A_PHI_BP2);
  int rc;
  // phi_bp = trial.psi (or trial.init) for every walker (back_propagation.py:150-153)
  ++h->launches;
  phi_to_of_kernel<<<grid_for((size_t)d.Wp * d.ne * d.Mp), 256, 0, st>>>(
      h->ptr<double2>(init_walker ? A_INIT_NAT : A_PSI_NAT), phi_bp, d, 1);
  PXB_CUDA(h, cudaGetLastError());
  // propagation/generic.py:277-287: configurations in reverse order, B^dagger each
  for (int i = 0; i < nsteps; ++i) {
    const int s = nsteps - 1 - i;
    ++h->launches;
    bp_field_kernel<<<grid_for((size_t)d.WG * d.NKC * 16), 256, 0, st>>>(h->ptr<double>(A_FC), h->ptr<double>(A_XF),
                                                                         d, h->nbp, s);
    PXB_CUDA(h, cudaGetLastError());
    if ((rc = run_vhs_gemm(h, st))) return rc;
    if ((rc = run_one_body(h, phi_bp, work, nullptr, st, true))) return rc;
    if ((rc = run_taylor(h, work, nullptr, st))) return rc;
    if ((rc = run_one_body(h, work, phi_bp, nullptr, st, true))) return rc;
    if (i != 0 && i % nstblz == 0) {
      if ((rc = run_qr(h, phi_bp, st))) return rc;
    }
  }
  // G_s = gab(phi_bp_s, phi_old_s)^T = conj(phi_bp_s) (phi_old_s^T conj(phi_bp_s))^-1 phi_old_s^T
  const int nmax = d.na > d.nb ? d.na : d.nb;
  const int nld = nmax | 1, nsq = nmax * nld;
  ++h->launches;
  bp_overlap_kernel<<<2 * d.Wp, 128, 0, st>>>(h->ptr<double>(A_PHI_OLD), phi_bp, h->ptr<double2>(A_OB), d, nld, nsq);
  PXB_CUDA(h, cudaGetLastError());
  const int nmt = (nmax + 7) >> 3;
  double* thbp = h->ptr<double>(A_THETA_BP);
  const double* pold = h->ptr<double>(A_PHI_OLD);
  rc = 1;
  if (nmt <= 1) rc = launch_theta<1>(h, pold, thbp, st);
  else if (nmt <= 2) rc = launch_theta<2>(h, pold, thbp, st);
  else if (nmt <= 3) rc = launch_theta<3>(h, pold, thbp, st);
  else if (nmt <= 4) rc = launch_theta<4>(h, pold, thbp, st);
  else if (nmt <= 5) rc = launch_theta<5>(h, pold, thbp, st);
  else if (nmt <= 6) rc = launch_theta<6>(h, pold, thbp, st);
  else if (nmt <= 8) rc = launch_theta<8>(h, pold, thbp, st);
  if (rc == 1) return fail(h, PXB_ERR_UNSUPPORTED, "back propagation: too many occupied orbitals for the Theta kernel");
  if (rc) return rc;
  // estimates += weight * G (back_propagation.py:198-205), BP-PhL weights
  BpRdmArgs a;
  a.phi_bp = phi_bp;
  a.theta = thbp;
  ++h->launches;
  bp_weight_kernel<<<(d.W + 255) / 256, 256, 0, st>>>(h->field<double>(PXB_F_WEIGHT), h->ptr0<double2>(A_BPFAC),
                                                     h->ptr0<double2>(A_BPW), h->bp_restore, d.W);
  PXB_CUDA(h, cudaGetLastError());
  a.weight = h->ptr0<double2>(A_BPW);
  a.part = h->ptr<double2>(A_BP_PART);
  a.d = d;
  a.nchunks = h->bp_chunks;
  a.wchunk = (d.W + a.nchunks - 1) / a.nchunks;
  a.tiles = (d.M + BPR_T - 1) / BPR_T;
  ++h->launches;
  bp_rdm_kernel<<<a.tiles * a.tiles * 2 * a.nchunks, 256, 0, st>>>(a);
  PXB_CUDA(h, cudaGetLastError());
  ++h->launches;
  bp_reduce_kernel<<<std::min(grid_for((size_t)2 * d.M * d.M), 4 * h->sm_count), 256, 0, st>>>(
      a.part, h->field<double2>(PXB_F_BP_RDM), h->field<double2>(PXB_F_BP_DENOM), a.weight, d, a.nchunks);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_bp_reset(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  if (h->nbp <= 0) return fail(h, PXB_ERR_STATE, "back propagation was not enabled (pxb_config.nbp == 0)");
  // copy_historic_wfn (walkers/handler.py:200-203) + FieldConfig.reset (walkers/stack.py:122-125)
  PXB_CUDA(h, cudaMemcpyAsync(h->ptr<void>(A_PHI_OLD), h->phi(), of_size(h->d) * 8, cudaMemcpyDeviceToDevice,
                              S(stream)));
  ++h->launches;
  bpfac_reset_kernel<<<(h->d.Wp + 255) / 256, 256, 0, S(stream)>>>(h->ptr0<double2>(A_BPFAC), h->d.Wp);
  PXB_CUDA(h, cudaGetLastError());
  h->bp_step = 0;
  return PXB_OK;
}

int pxb_bp_restore_weights(pxb_handle h, int mode) {
  if (!h || mode < 0 || mode > 2) return PXB_ERR_ARG;
  h->bp_restore = mode;
  return PXB_OK;
}

int pxb_bp_zero(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  if (h->nbp <= 0) return fail(h, PXB_ERR_STATE, "back propagation was not enabled (pxb_config.nbp == 0)");
  PXB_CUDA(h, cudaMemsetAsync(h->field<void>(PXB_F_BP_RDM), 0, (size_t)2 * h->d.M * h->d.M * 16, S(stream)));
  PXB_CUDA(h, cudaMemsetAsync(h->field<void>(PXB_F_BP_DENOM), 0, 16, S(stream)));
  return PXB_OK;
}

int pxb_get_phi_bp(pxb_handle h, int which, void* dev_out, void* stream) {
  PXB_REQUIRE_READY(h);
  if (h->nbp <= 0) return fail(h, PXB_ERR_STATE, "back propagation was not enabled (pxb_config.nbp == 0)");
  const Dims& d = h->d;
  ++h->launches;
  of_to_natural_kernel<<<grid_for((size_t)d.W * d.ne * d.M), 256, 0, S(stream)>>>(
      h->ptr<double>(which ? A_PHI_OLD : A_PHI_BP), static_cast<double2*>(dev_out), d, 1);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

// ---- stage-level entry points -------------------------------------------------
int pxb_stage_exchange(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  int rc = ensure_theta(h, S(stream));
  if (rc) return rc;
  return run_exchange(h, S(stream));
}

int pxb_stage_greens(pxb_handle h, int with_e1b, void* stream) {
  PXB_REQUIRE_READY(h);
  (void)with_e1b;
  h->theta_valid = h->x_valid = h->eloc_valid = false;
  return ensure_theta(h, S(stream));
}

int pxb_stage_force_bias_gemm(pxb_handle h, void* stream) {
  PXB_REQUIRE_READY(h);
  int rc = ensure_theta(h, S(stream));
  if (rc) return rc;
  h->x_valid = false;
  return ensure_x(h, S(stream));
}

int pxb_get_theta(pxb_handle h, void* dev_theta, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  ++h->launches;
  of_to_natural_kernel<<<grid_for((size_t)d.W * d.ne * d.M), 256, 0, S(stream)>>>(
      h->ptr<double>(A_THETA), static_cast<double2*>(dev_theta), d, 0);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_get_x(pxb_handle h, void* dev_x, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  // X is stored [2][Wp][Np]; compact to [2][W][N]
  for (int s = 0; s < 2; ++s)
    PXB_CUDA(h, cudaMemcpy2DAsync(static_cast<char*>(dev_x) + (size_t)s * d.W * d.N * 16, (size_t)d.N * 16,
                                  h->ptr<char>(A_X) + (size_t)s * d.Wp * d.Np * 16, (size_t)d.Np * 16,
                                  (size_t)d.N * 16, d.W, cudaMemcpyDeviceToDevice, S(stream)));
  return PXB_OK;
}

int pxb_get_vhs(pxb_handle h, void* dev_vhs, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  ++h->launches;
  vf_to_natural_kernel<<<grid_for((size_t)d.W * d.M * d.M), 256, 0, S(stream)>>>(
      h->ptr<double>(A_VF), static_cast<double2*>(dev_vhs), d);
  PXB_CUDA(h, cudaGetLastError());
  return PXB_OK;
}

int pxb_get_exx(pxb_handle h, void* dev_exx, void* stream) {
  PXB_REQUIRE_READY(h);
  const Dims& d = h->d;
  for (int s = 0; s < 2; ++s)
    PXB_CUDA(h, cudaMemcpyAsync(static_cast<char*>(dev_exx) + (size_t)s * d.W * 16,
                                h->ptr<char>(A_EXX) + (size_t)s * d.Wp * 16, (size_t)d.W * 16,
                                cudaMemcpyDeviceToDevice, S(stream)));
  return PXB_OK;
}

}  // extern "C"

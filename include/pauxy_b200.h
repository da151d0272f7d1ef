/*
 * pauxy_b200 -- C-ABI of the B200 phaseless-AFQMC walker-propagation hot path.
 *
 * The reference (pauxy-qmc/pauxy) is pure Python and has no FFI: its plugin
 * surface for this path is the duck-typed propagator / walker / estimator
 * classes called from AFQMC.run (pauxy/qmc/afqmc.py:223-255).  This header is
 * the boundary a ctypes binding in those classes would call; each entry point
 * names the reference code it replaces.  Plain pointers and sizes only.
 *
 * Conventions
 *   - every function returns 0 on success, a negative PXB_ERR_* otherwise;
 *     pxb_last_error() returns a message.  Nothing calls exit().
 *   - "dev" pointers are CUDA device pointers (e.g. torch.Tensor.data_ptr()).
 *   - complex128 arrays are interleaved (re, im) doubles, C order, in the
 *     reference's own layouts.
 *   - all work is enqueued on the caller's cudaStream_t (passed as void*);
 *     calls are asynchronous unless stated otherwise.
 */
#ifndef PAUXY_B200_H
#define PAUXY_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PXB_OK 0
#define PXB_ERR_ARG -1          /* bad argument / shape */
#define PXB_ERR_CUDA -2         /* CUDA runtime error (see pxb_last_error) */
#define PXB_ERR_STATE -3        /* call order (arena / hamiltonian not set) */
#define PXB_ERR_UNSUPPORTED -4  /* e.g. complex-valued Cholesky / trial */

#define PXB_ABI_VERSION 10

typedef struct pxb_context* pxb_handle;

/* Problem description.  Mirrors system.nbasis/nup/ndown/nfields
 * (pauxy/systems/generic.py:74-166), qmc.dt (pauxy/qmc/options.py:88), and the
 * propagator option expansion_order (pauxy/propagation/continuous.py:37). */
typedef struct {
  int32_t nbasis;     /* M */
  int32_t nup;        /* na */
  int32_t ndown;      /* nb */
  int32_t nchol;      /* N = nfields */
  int32_t nwalkers;   /* walkers on THIS device */
  int32_t exp_order;  /* Taylor order of exp(VHS), reference default 6 */
  int32_t device;     /* CUDA device ordinal */
  int32_t total_walkers; /* walkers over ALL devices (0: == nwalkers) */
  double dt;
  int32_t exchange_mode; /* PXB_EXCHANGE_*: how pxb_local_energy evaluates the exchange term */
  int32_t flags;         /* PXB_FLAG_* */
  int32_t nbp;           /* back propagation: field configurations kept per walker
                            (estimators/back_propagation.py:55 nmax = int(tau_bp/dt)); 0 = off */
  int32_t ndets;         /* determinants of the trial (walkers/multi_det.py); 0 or 1: single determinant.
                            The per-determinant operands are supplied with pxb_set_trial_det. */
} pxb_config;
#define PXB_MAX_DETS 8

/* Shape limits of this version (a call outside them returns PXB_ERR_ARG / PXB_ERR_UNSUPPORTED with a
 * message in pxb_last_error; nothing is evaluated by a slower path silently):
 *   nbasis <= 256                        Taylor kernels: taylor3 4 <= ceil(M/8) <= 16, taylor2 up to M = 224 with at
 *                                        most 48 orbitals per item, the direct kernel up to 256
 *   nup, ndown <= 64                     Green's function and exchange kernels; up to 32 per spin the inverse and
 *                                        the re-orthogonalisation run register-resident (Gauss-Jordan, CholeskyQR2),
 *                                        above in shared memory (Gauss-Jordan, modified Gram-Schmidt)
 *   ndets <= PXB_MAX_DETS                and no back propagation with several determinants
 *   half-rotated ERI (exchange_mode ERI / AUTO): (nup^2 + ndown^2) M^2 doubles per determinant; AUTO falls
 *                                        back to the Cholesky form above 16 GiB
 *   the overlap / Theta / QR tiles of one (walker, spin) must fit 227 KB of shared memory (true within the
 *   limits above). */

/* propagator options of pauxy/propagation/continuous.py:14-33 */
#define PXB_FLAG_FREE_PROJECTION 1 /* propagate_walker_free (continuous.py:175-200), the free-projection
                                      branches of Walkers.orthogonalise (handler.py:178-181) and of
                                      Mixed.update (mixed.py:151-177); implies NO_FORCE_BIAS like the
                                      reference (continuous.py:30-33) */
#define PXB_FLAG_NO_FORCE_BIAS 2   /* force_bias = False: xbar = 0 (continuous.py:136-138) */
#define PXB_FLAG_LOCAL_ENERGY_WEIGHT 4 /* propagator.hybrid = false: update_weight_local_energy
                                      (continuous.py:216-231,294-318) instead of the hybrid update.  The
                                      reference supports it with MultiDetWalker only (its SingleDet +
                                      Generic path raises TypeError); here it is available for any
                                      number of determinants with the multi-determinant semantics. */

#define PXB_FLAG_COMPLEX_CHOLESKY 16 /* complex-valued Cholesky vectors (system.hs_pot, trial._rchol) and / or trial
                                       orbitals: every GEMM against them runs as the real GEMM
                                       [Re A | Im A] [B ; i B] over a doubled k range, the exchange as two
                                       quadratic forms (Re K, Im K).  ERI form of the exchange only; no back
                                       propagation. */
#define PXB_FLAG_COMPLEX_ONE_BODY 8 /* propagator.BH1 has an imaginary part (complex mean-field shift, e.g.
                                      a multi-determinant trial with complex CI coefficients): the
                                      one-body step runs as [Re BH1 | Im BH1] [phi ; i phi] on the real
                                      GEMM kernel; without the flag a complex BH1 is PXB_ERR_UNSUPPORTED */

/* Exchange energy of estimators/generic.py:198-214.  Both forms give the same number to
 * rounding (the ERI is rebuilt from the same Cholesky vectors):
 *   CHOLESKY  T[x] = R[x] Theta^T per Cholesky vector, fused trace   4 N ns^2 M flop / walker / spin
 *   ERI       quadratic form Theta . K . Theta in the half-rotated ERI (the contraction of
 *             local_energy_generic_opt, estimators/generic.py:133-150), K symmetric,
 *             2 (ns M)^2 flop / walker / spin, K = (ns M)^2 doubles per spin in the arena
 *   AUTO      ERI when K (both spins) fits in 16 GiB, else CHOLESKY */
#define PXB_EXCHANGE_AUTO 0
#define PXB_EXCHANGE_CHOLESKY 1
#define PXB_EXCHANGE_ERI 2

/* Walker-state fields living in the arena; pxb_field() gives their location
 * so the host can view them (e.g. as torch tensors over the arena).  They are
 * the per-walker scalars of pauxy/walkers/walker.py:24-61. */
enum pxb_field_id {
  PXB_F_WEIGHT = 0,          /* f64  [W]    walker.weight              */
  PXB_F_UNSCALED_WEIGHT = 1, /* f64  [W]    walker.unscaled_weight     */
  PXB_F_OT = 2,              /* c128 [W]    walker.ot == walker.ovlp   */
  PXB_F_HYBRID_ENERGY = 3,   /* c128 [W]    walker.hybrid_energy       */
  PXB_F_ELOC = 4,            /* c128 [W,3]  (E, E1, E2) last evaluated */
  PXB_F_DETR = 5,            /* f64  [W]    walker.detR                */
  PXB_F_LOG_DETR = 6,        /* f64  [W]    walker.log_detR            */
  PXB_F_ESTIMATES = 7,       /* c128 [10]   Mixed.estimates accumulators (mixed.py:460-469) */
  PXB_F_COUNTERS = 8,        /* i64  [8]    nfb_trig, nhe_trig, n_inactive, n_comb_moves, vanished (sticky: total weight < 1e-8 seen, handler.py:236-241) */
  PXB_F_PARENT_IX = 9,       /* i32  [Wtot] comb parent_ix of the last pop-control */
  PXB_F_XBAR = 10,           /* c128 [W,N]  force bias after clipping (debug/parity) */
  PXB_F_XSHIFTED = 11,       /* c128 [W,N]  x = xi - xbar (natural layout copy, debug/parity) */
  PXB_F_CMF_CFB = 12,        /* c128 [W,2]  (cmf, cfb) of the last propagate */
  PXB_F_OVLP_NEW = 13,       /* c128 [W]    overlap after the last propagate */
  PXB_F_TOTAL_WEIGHT = 14,   /* f64  [1]    walker.total_weight (same for all walkers) */
  PXB_F_PAIRS = 15,          /* i32  [1+2*Wtot] n_pairs then (clone, kill) global indices */
  PXB_F_PHASE = 16,          /* c128 [W]    walker.phase (free projection; 1 otherwise) */
  PXB_F_BP_RDM = 17,         /* c128 [2,M,M] sum_w weight_w G_w of the back-propagated estimator
                                              (back_propagation.py:198-205), nbp > 0 only */
  PXB_F_BP_DENOM = 18,       /* c128 [1]    sum_w weight_w (back_propagation.py:200) */
  PXB_F_THETA_SUM = 19,      /* c128 [ne,M] sum_w weight_w Theta_w since the last pxb_zero_estimates: the mixed
                                one-body density matrix is Re(conj(psi) THETA_SUM) (mixed.py:226-229) */
  PXB_F_WALKER_ELOC = 20,    /* c128 [W]    walker.eloc of the local-energy weight update (walker.py:37) */
  PXB_F_OVLP_DET = 21,       /* c128 [ndets, W(padded to 4, row stride a multiple of 256 bytes)] overlaps of
                                the walkers with the single determinants of the trial (MultiDetWalker.ovlps) */
  PXB_F_LOG_SHIFTS = 22,     /* walkers.use_log_shift: f64 log_shift, detR_shift, log_detR_shift, i64 shift
                                counter, ... ; at byte 64: f64 [3] population sums |ot|, |detR|, |log_detR| of
                                this device (all-reduce them between pxb_log_shift_sums and _update) */
  PXB_F_COUNT = 23
};

int pxb_abi_version(void);
/* number of CUDA kernels launched through this handle so far (bench.py gpu_launches) */
long long pxb_launch_count(pxb_handle h);

/* Optional per-stage timing: with profiling enabled every stage of the hot path is bracketed by
 * CUDA events on the stream it is launched on.  pxb_stage_times waits for the recorded events
 * and returns the accumulated milliseconds / call counts per PXB_STAGE_* (arrays of n entries). */
enum pxb_stage_id {
  PXB_STAGE_GREENS = 0,     /* overlap, inverse, Theta, e1b          (single_det.py:295-321)   */
  PXB_STAGE_XGEMM = 1,      /* X = R^T Theta, force bias == Coulomb  (generic.py:130-152)      */
  PXB_STAGE_FIELD = 2,      /* xbar clip, x = xi - xbar, cmf, cfb    (continuous.py:133-158)   */
  PXB_STAGE_VHS = 3,        /* VHS = i sqrt(dt) L x                  (generic.py:164-179)      */
  PXB_STAGE_ONE_BODY = 4,   /* phi = BH1 phi, both half steps        (operations.py:29-52)     */
  PXB_STAGE_TAYLOR = 5,     /* exp(VHS) phi                          (continuous.py:82-111)    */
  PXB_STAGE_WEIGHT = 6,     /* hybrid weight update + cap            (continuous.py:264-292)   */
  PXB_STAGE_EXCHANGE = 7,   /* exchange energy                       (estimators/generic.py:198-214) */
  PXB_STAGE_ENERGY = 8,     /* Coulomb + assembly of (E, E1, E2)     (estimators/generic.py:187-221) */
  PXB_STAGE_QR = 9,         /* re-orthogonalisation                  (single_det.py:215-255)   */
  PXB_STAGE_POP_CONTROL = 10, /* comb on one device                  (handler.py:225-338)      */
  PXB_STAGE_ACCUMULATE = 11,  /* Mixed.update sums                   (mixed.py:211-225)        */
  PXB_STAGE_COUNT = 12
};
int pxb_exchange_mode(pxb_handle h); /* PXB_EXCHANGE_CHOLESKY or PXB_EXCHANGE_ERI in effect */
/* 1 if system.hs_pot was found symmetric in (p,q) (real orbitals): the VHS GEMM then computes the
 * upper triangle only and mirrors it; 0 otherwise (after pxb_set_hamiltonian) */
int pxb_vhs_symmetric(pxb_handle h);
int pxb_profile(pxb_handle h, int enable);
int pxb_stage_times(pxb_handle h, double* ms, long long* calls, int n, int reset);

/* ---- lifecycle -------------------------------------------------------- */
int pxb_create(pxb_handle* out, const pxb_config* cfg);
int pxb_destroy(pxb_handle h);
const char* pxb_last_error(pxb_handle h);

/* Device memory is owned by the caller: one arena of pxb_arena_bytes() bytes
 * (256-byte aligned).  pxb_bind_arena zero-fills it on `stream`. */
int pxb_arena_bytes(pxb_handle h, size_t* bytes);
int pxb_bind_arena(pxb_handle h, void* dev_arena, size_t bytes, void* stream);
int pxb_field(pxb_handle h, int field_id, size_t* offset_bytes, size_t* size_bytes);

/* ---- Hamiltonian / trial (setup, once) ----------------------------------
 * Device pointers to the reference's arrays, read during the call only:
 *   hs_pot   f64  [M*M, N]      system.hs_pot            (systems/generic.py:154)
 *   rchol    c128 [(na+nb)*M,N] trial._rchol             (trial_wavefunction/multi_slater.py:370-418)
 *   bh1      c128 [2, M, M]     propagator.BH1           (propagation/generic.py:106-107)
 *   h1rot    c128 [(na+nb), M]  psi_s^dagger H1[s] (rows: up orbitals, then down);
 *                               e1b = sum h1rot * Theta == sum H1*G (estimators/generic.py:178)
 *   psi      c128 [M, na+nb]    trial.psi                (walkers/handler.py:57-61)
 *   mf_shift c128 [N]           propagator.mf_shift      (propagation/generic.py:66-80)
 * Without PXB_FLAG_COMPLEX_CHOLESKY rchol and psi must be real-valued (imag == 0) and hs_pot is an
 * array of doubles; PXB_ERR_UNSUPPORTED otherwise (checked on the device, synchronous call).  With
 * the flag (systems/generic.py:126 "complex integrals", generate_hamiltonian(cplx=True)) hs_pot is
 * complex128 [M*M, N] (interleaved re, im: pass the pointer as const double*), rchol and psi may be
 * complex.  A complex bh1 needs PXB_FLAG_COMPLEX_ONE_BODY in both cases. */
int pxb_set_hamiltonian(pxb_handle h, const double* dev_hs_pot, const void* dev_rchol,
                        const void* dev_bh1, const void* dev_h1rot, const void* dev_psi,
                        const void* dev_mf_shift, double ecore, void* stream);

/* Multi-determinant trial (walkers/multi_det.py:27-300, propagation/generic.py:154-157,
 * estimators/mixed.py:439-448; pxb_config.ndets > 1).  pxb_set_hamiltonian supplies determinant 0
 * (and the determinant-independent arrays); this call supplies determinant `det` >= 1 -- its
 * orbitals, half-rotated Cholesky vectors and half-rotated one-body integrals, same layouts as
 * above -- and the CI coefficient of ANY determinant (det 0 included: pass NULL arrays to set the
 * coefficient only; it defaults to 1).  The overlap is sum_i conj(c_i) <psi_i|phi>; force bias and
 * local energy are the weighted averages over the determinants with w_i = conj(c_i) <psi_i|phi>.
 * The orbitals must be real-valued like the single-determinant ones; the coefficients may be
 * complex.  mf_shift / bh1 passed to pxb_set_hamiltonian are the multi-determinant ones
 * (propagation/generic.py:82-86). */
int pxb_set_trial_det(pxb_handle h, int det, double coeff_re, double coeff_im, const void* dev_rchol,
                      const void* dev_h1rot, const void* dev_psi, void* stream);

/* Imaginary part of the energy shift used by the following pxb_propagate / pxb_step calls (zero
 * by default and in the driver; the reference's own propagation tests pass the complex trial
 * energy, propagation/tests/test_generic.py:66-68). */
int pxb_set_eshift_imag(pxb_handle h, double eshift_im);

/* ---- walker state --------------------------------------------------------
 * phi: c128 [W, M, na+nb] device, the reference layout of walker.phi stacked
 * over walkers (walkers/walker.py:29). */
int pxb_set_phi(pxb_handle h, const void* dev_phi, void* stream);
int pxb_get_phi(pxb_handle h, void* dev_phi, void* stream);
/* Walkers.__init__ + SingleDetWalker.__init__ (walkers/handler.py:36-164,
 * walkers/single_det.py:31-94): phi = trial.init for every walker, weight = 1,
 * ot = calc_overlap, hybrid_energy = 0, total_weight = total_walkers. */
int pxb_init_walkers(pxb_handle h, const void* dev_init_phi /* c128 [M,ne] */,
                     double total_walkers, void* stream);

/* ---- the hot path -------------------------------------------------------- */
/* Continuous.propagate_walker_phaseless for every walker with |weight| > 1e-8,
 * followed by the 10 % weight cap (propagation/continuous.py:232-292,
 * qmc/afqmc.py:231-236).
 *   dev_xi : f64 [W, N] auxiliary fields xi for this step (row w is read only
 *            if walker w is active), or NULL to draw them on the device with
 *            Philox4x32-10 keyed by (rng_seed, step, global walker index).
 *   walker_offset : global index of local walker 0 (rank * W), Philox only. */
int pxb_propagate(pxb_handle h, const double* dev_xi, uint64_t rng_seed,
                  int64_t walker_offset, double eshift, int64_t step, void* stream);

/* Walkers.orthogonalise -> SingleDetWalker.reortho (walkers/handler.py:166-181,
 * walkers/single_det.py:215-255), phaseless branch. */
int pxb_orthogonalise(pxb_handle h, void* stream);

/* One whole pass of the driver loop body (qmc/afqmc.py:223-255) on ONE device as a single call:
 * [pxb_orthogonalise] -> pxb_propagate -> [comb population control, its plan overlapped with
 * pxb_local_energy] -> pxb_accumulate.  Same results as the separate calls; the launch sequence is
 * captured as a CUDA graph the second time a (flags, dev_xi, state) variant is seen and replayed
 * afterwards (small problems are bound by launch latency).  The per-step scalars travel through a
 * device-side parameter block, not through kernel arguments.  comb_r: the comb's uniform draw
 * (walkers/handler.py:275).  Not available with back propagation (nbp > 0) or several devices. */
enum pxb_step_flags {
  PXB_STEP_ORTHO = 1,  /* re-orthogonalise first (walkers/handler.py:166-181)            */
  PXB_STEP_POP = 2,    /* comb population control after the propagation (handler.py:225-338) */
  PXB_STEP_ENERGY = 4  /* local energy of every walker, accumulated with the other sums  */
};
int pxb_step(pxb_handle h, const double* dev_xi, uint64_t rng_seed, int64_t walker_offset,
             double eshift, int64_t step, double comb_r, int flags, void* stream);
/* enable (1) / disable (0) / query (-1) graph replay in pxb_step; *replays = graph launches so far */
int pxb_step_graphs(pxb_handle h, int enable, long long* replays);

/* greens_function + local_energy_generic_cholesky_opt for every walker
 * (walkers/single_det.py:295-321, estimators/generic.py:156-221) -> ELOC. */
int pxb_local_energy(pxb_handle h, void* stream);

/* Mixed.update accumulation (estimators/mixed.py:211-225) into ESTIMATES;
 * with_energy != 0 adds the enumer/e1b/e2b/edenom terms from ELOC. */
int pxb_accumulate(pxb_handle h, int with_energy, void* stream);
int pxb_zero_estimates(pxb_handle h, void* stream);
/* Mixed.update with one_rdm (estimators/mixed.py:226-229): THETA_SUM += sum_w weight_w Theta_w */
int pxb_accumulate_theta(pxb_handle h, void* stream);

/* ---- population control (walkers/handler.py:225-412) -------------------- */
/* Single-device pop_control with the comb: total weight (sequential sum),
 * rescale, comb selection with the caller's uniform r (numpy.random.random()
 * in the reference, handler.py:276), walker copies, weights reset to 1.
 * Entirely on the device, no host synchronisation. */
/* walkers.use_log_shift (walkers/handler.py:228,456-475): running population averages that rescale
 * the stored overlaps and detR factors.  pop_control does: pxb_log_shift_sums -> all-reduce of the
 * three sums over the devices (the caller's collective) -> pxb_log_shift_update.  The walk itself
 * (weights, energies, selection) is unaffected: every ratio the propagation uses is shift-free. */
int pxb_log_shift_enable(pxb_handle h, int enable, void* stream);
int pxb_log_shift_sums(pxb_handle h, void* stream);
int pxb_log_shift_update(pxb_handle h, void* stream);

int pxb_pop_control_comb(pxb_handle h, double r, void* stream);

/* Multi-device pieces.  dev_global_abs_weights: f64 [Wtot] = |weight| of all
 * walkers in global order (all-gathered by the caller).
 *  pxb_pop_rescale : total = sequential sum; unscaled_weight = weight;
 *                    weight /= total/Wtot; TOTAL_WEIGHT = total
 *                    (handler.py:233-249).
 *  pxb_comb_plan   : parent_ix (PARENT_IX) and the (clone, kill) pair list
 *                    (PAIRS) from the rescaled global weights (handler.py:271-301).
 *  pxb_pair_branch_plan_host : host-side pair_branch selection (handler.py:
 *                    340-386) -- sorting + uniform draws happen on the host. */
int pxb_pop_rescale(pxb_handle h, const double* dev_global_abs_weights, int64_t wtot,
                    void* stream);
int pxb_comb_plan(pxb_handle h, const double* dev_global_abs_weights, int64_t wtot, double r,
                  void* stream);
/* Walker payload movement (Walker.get_buffer/set_buffer, walkers/walker.py:
 * 63-131): the fields that matter downstream -- phi, weight, unscaled_weight,
 * ot, hybrid_energy, phase, eloc, detR, log_detR -- plus what this library keeps
 * per walker so that nothing has to be recomputed after a copy: the rotated
 * Green's function Theta with its one-body energy, X = R^T Theta, and with
 * nbp > 0 phi_old and the field history (walker.field_configs).
 *  pxb_copy_walkers : local slot src[i] -> local slot dst[i], i < n (int32 dev lists)
 *  pxb_pack_walkers / pxb_unpack_walkers : to / from a contiguous buffer of
 *      n * pxb_payload_doubles() doubles, for send/recv between devices. */
int pxb_payload_doubles(pxb_handle h, size_t* ndoubles);
int pxb_copy_walkers(pxb_handle h, const int32_t* dev_src, const int32_t* dev_dst, int n,
                     void* stream);
int pxb_pack_walkers(pxb_handle h, const int32_t* dev_slots, int n, double* dev_buffer,
                     void* stream);
int pxb_unpack_walkers(pxb_handle h, const int32_t* dev_slots, int n, const double* dev_buffer,
                       void* stream);
int pxb_set_weights(pxb_handle h, double value, void* stream); /* handler.py:337-338 */

/* Peer-memory comb for several devices of one node (walkers/handler.py:225-338 with the
 * Isend/Recv of handler.py:301-334 replaced by direct NVLink reads).  The reference moves a cloned
 * walker with MPI point-to-point messages whose sizes the host has to know; here every device maps
 * the arenas of its peers (CUDA IPC) and pulls the clones of its own killed walkers with one
 * kernel, so a pop-control step needs no host synchronisation at all.
 *  pxb_peer_export : IPC handle (PXB_IPC_HANDLE_BYTES bytes) of the cudaMalloc block holding the
 *                    arena and the arena's offset inside it; exchange them between the ranks
 *                    (e.g. an all-gather), then
 *  pxb_peer_attach : map every peer's arena.  handles: [nranks][PXB_IPC_HANDLE_BYTES] bytes,
 *                    offsets: [nranks].  All ranks must use the same pxb_config except `device`.
 *  pxb_pop_control_comb_peers : total weight, comb plan (identical on every rank from the
 *                    all-gathered |weights|, r = the caller's uniform) and the pull of the clones.
 *                    The caller must run a cross-device barrier ON THE STREAM (e.g. a 1-element
 *                    all-reduce) between this call and
 *  pxb_pop_control_finish : unscaled_weight = weight, weight = 1 (handler.py:247-248, :337-338);
 *                    nothing that writes walker state may be enqueued before that barrier. */
/* The same, split so that the host can overlap the plan with the local-energy evaluation (the plan
 * needs only the weights; the energies of the walkers before the comb are the energies of their
 * clones after it, and ELOC / X / Theta travel in the walker payload):
 *   side stream  : [all-gather |w|]  pxb_pop_plan
 *   launch stream: pxb_local_energy   ... wait for the side stream ...
 *                  [stream barrier across devices: the peers' ELOC / X are final]  pxb_pop_pull
 *                  [stream barrier across devices: the peers have finished reading]  pxb_pop_control_finish
 * pxb_pop_plan    : total weight + comb plan, writes no walker state; dev_global_abs_weights may be
 *                   NULL on one device.  pxb_pop_pull: the data movement of that plan (local
 *                   copies, NVLink pulls for clones owned by peers; on one device no attach needed). */
int pxb_pop_plan(pxb_handle h, const double* dev_global_abs_weights, int64_t wtot, double r,
                 void* stream);
int pxb_pop_pull(pxb_handle h, void* stream);
/* The persistent kernels of pxb_local_energy (X GEMM, exchange) normally take one CTA per SM, which
 * leaves no room for a side-stream kernel until they end.  pxb_reserve_sms(h, n) makes them leave n
 * SMs free (n = 1 while a long comb plan -- tens of thousands of walkers over all devices -- runs
 * beside them; 0 restores the default). */
int pxb_reserve_sms(pxb_handle h, int n);

#define PXB_IPC_HANDLE_BYTES 64
int pxb_peer_export(pxb_handle h, void* handle_out, uint64_t* offset_out);
int pxb_peer_attach(pxb_handle h, int rank, int nranks, const void* handles,
                    const uint64_t* offsets);
int pxb_pop_control_comb_peers(pxb_handle h, const double* dev_global_abs_weights, int64_t wtot,
                               double r, void* stream);
int pxb_pop_control_finish(pxb_handle h, void* stream);

/* ---- back propagation (pxb_config.nbp > 0) ------------------
 * pxb_propagate then keeps the shifted fields x of every step per walker (FieldConfig.update,
 * walkers/stack.py:52-79 called from propagation/continuous.py:288-289); they and phi_old travel
 * with the walker through population control.
 *  pxb_bp_steps        : configurations stored since the last reset (field_configs.step).
 *  pxb_back_propagate  : BackPropagation.update_uhf (estimators/back_propagation.py:127-225) for
 *      all walkers: phi_bp = trial.psi (init_walker != 0: trial.init), then for the first `nsteps`
 *      stored configurations in reverse order phi_bp <- B(c)^dagger phi_bp
 *      (back_propagate_generic, propagation/generic.py:253-290; B of generic.py:180-213) with a QR
 *      re-orthogonalisation after every nstblz applications, G = gab(phi_bp, phi_old)^T and
 *      BP_RDM += weight * G, BP_DENOM += weight (BP-PhL weights).  Needs Cholesky matrices that are
 *      symmetric to rounding (real orbitals): PXB_ERR_UNSUPPORTED otherwise; BH1 may be any real matrix.
 *  pxb_bp_reset        : phi_old = phi for every walker, empty history (walkers/handler.py:200-203,
 *      walkers/stack.py:122-125).
 *  pxb_bp_zero         : BackPropagation.zero (back_propagation.py:335-338).
 *  pxb_get_phi_bp      : c128 [W, M, ne] copy of the back-propagated (which = 0) or the historic
 *      (which = 1) determinants (tests). */
int pxb_bp_steps(pxb_handle h);
int pxb_back_propagate(pxb_handle h, int nsteps, int nstblz, int init_walker, void* stream);
int pxb_bp_reset(pxb_handle h, void* stream);
/* estimator weights of the following pxb_back_propagate calls (back_propagation.py:75-80,187-196):
 * 0 BP-PhL (walker.weight), 1 restore_weights = "partial" (times the product of the phase factors
 * I/|I| of the stored steps), 2 "full" (also divided by the product of the cosine factors). */
int pxb_bp_restore_weights(pxb_handle h, int mode);
int pxb_bp_zero(pxb_handle h, void* stream);
int pxb_get_phi_bp(pxb_handle h, int which, void* dev_out, void* stream);

/* Host helpers (pure C, no device): bit-exact restatements used by the host
 * mirror when the selection has to happen on the host. */
int pxb_comb_plan_host(const double* weights, int64_t n, double r, int32_t* parent_ix);

/* ---- stage-level entry points (tests, profiling) -------------------------
 * Each runs one stage of the pipeline on the current state. */
int pxb_stage_greens(pxb_handle h, int with_e1b, void* stream);      /* A1: Theta, ovlp */
int pxb_stage_force_bias_gemm(pxb_handle h, void* stream);           /* A3/C1: X_s = R_s^T Theta_s */
int pxb_stage_exchange(pxb_handle h, void* stream);                  /* C1: exx_s from the current Theta */
int pxb_get_theta(pxb_handle h, void* dev_theta /* c128 [W, ne, M] */, void* stream);
int pxb_get_x(pxb_handle h, void* dev_x /* c128 [2, W, N] */, void* stream);
int pxb_get_vhs(pxb_handle h, void* dev_vhs /* c128 [W, M, M] */, void* stream);
int pxb_get_exx(pxb_handle h, void* dev_exx /* c128 [2, W] */, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* PAUXY_B200_H */

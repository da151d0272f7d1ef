"""TEST INFRASTRUCTURE -- CPU (numpy) restatement of the reference hot path.

Phaseless AFQMC propagation + local energy for a generic Cholesky Hamiltonian
with a single-determinant trial, restated from the reference files cited on
each function (paths under /root/reference/).  It is the checker that travels
to the GPU box (the reference itself cannot).  Only tests/, __graft_entry__.
smoke() and bench.py's cpu_baseline / --impl reference legs may import it; the
product path (pauxy_b200/) never does.

PARITY PINNED: tests/test_oracle_cpu.py checks this module against
  * the reference's own driver goldens (pauxy/qmc/tests/test_afqmc.py:227,229)
  * the reference's local-energy golden (pauxy/estimators/tests/test_generic.py:45-47)
  * per-step traces of the unmodified reference recorded by oracle/gen_golden.py
    into tests/golden/*.npz (weights, overlaps, hybrid energies, local energies,
    comb parent indices, estimator rows).

Arithmetic is IEEE double on complex128 arrays as in the reference; walkers are
looped or batched with numpy.matmul (sum order differs from the reference's
per-walker BLAS calls at the 1e-16 level only).
"""
import cmath
import math

import numpy
import scipy.linalg


# --------------------------------------------------------------------------
# setup (host side in the reference as well)
# --------------------------------------------------------------------------
class Hamiltonian(object):
    """Arrays the reference derives at construction.

    systems/generic.py:74-166,202-210 (h1e_mod), trial_wavefunction/utils.py:
    64-77 (RHF guess), trial_wavefunction/multi_slater.py:60-63,267-420 (G, half
    rotation), propagation/generic.py:29-107 (mean-field shift, BH1, mf_core).
    """

    def __init__(self, h1e, hs_pot, ecore, nelec, dt, psi=None):
        h1e = numpy.asarray(h1e)
        if h1e.ndim == 2:
            h1e = numpy.array([h1e, h1e])
        self.H1 = h1e
        self.hs_pot = numpy.ascontiguousarray(hs_pot)
        self.ecore = ecore
        self.nup, self.ndown = nelec
        self.ne = self.nup + self.ndown
        M = h1e.shape[-1]
        self.nbasis = M
        self.nchol = self.hs_pot.shape[1]
        self.dt = dt
        self.sqrt_dt = dt ** 0.5
        self.isqrt_dt = 1j * self.sqrt_dt
        self.ebound = (2.0 / dt) ** 0.5
        na, nb = self.nup, self.ndown
        if psi is None:
            # trial_wavefunction/utils.py:64-74
            psi = numpy.zeros((M, na + nb), dtype=numpy.complex128)
            I = numpy.identity(M, dtype=numpy.complex128)
            psi[:, :na] = I[:, :na]
            psi[:, na:] = I[:, :nb]
        self.psi = psi
        # systems/generic.py:202-210
        chol3 = self.hs_pot.reshape((M, M, -1))
        v0 = 0.5 * numpy.einsum('ikn,jkn->ij', chol3, chol3, optimize='optimal')
        self.h1e_mod = numpy.array([h1e[0] - v0, h1e[1] - v0])
        # estimators/greens_function.py:41-79 (gab_mod with A = B = psi)
        G = []
        for s, sl in enumerate((slice(0, na), slice(na, na + nb))):
            A = psi[:, sl]
            O = numpy.dot(A.T, A.conj())
            Gh = numpy.dot(scipy.linalg.inv(O), A.T)
            G.append(numpy.dot(A.conj(), Gh))
        self.G = numpy.array(G)
        # multi_slater.py:402-409
        rup = numpy.tensordot(psi[:, :na].conj(), chol3, axes=((0), (0))).reshape((na * M, -1))
        rdn = numpy.tensordot(psi[:, na:].conj(), chol3, axes=((0), (0))).reshape((nb * M, -1))
        self.rchol = numpy.concatenate([rup, rdn]).astype(numpy.complex128)
        # propagation/generic.py:66-80
        self.mf_shift = 1j * numpy.dot(self.hs_pot.T, (self.G[0] + self.G[1]).ravel())
        # propagation/generic.py:88-107
        shift = 1j * self.hs_pot.dot(self.mf_shift).reshape(M, M)
        H1 = self.h1e_mod - numpy.array([shift, shift])
        self.BH1 = numpy.array([scipy.linalg.expm(-0.5 * dt * H1[0]),
                                scipy.linalg.expm(-0.5 * dt * H1[1])])
        self.mf_core = ecore + 0.5 * numpy.dot(self.mf_shift, self.mf_shift)


# --------------------------------------------------------------------------
# per-walker stage functions (batched over a leading walker axis)
# --------------------------------------------------------------------------
def _slogdet_prod(O):
    sign, logdet = numpy.linalg.slogdet(O)
    return sign, logdet


def greens_function(ham, phi):
    """walkers/single_det.py:295-321.  phi [W,M,ne] -> (theta_up [W,na,M],
    theta_dn [W,nb,M], det [W])."""
    na = ham.nup
    psi = ham.psi
    Oa = numpy.matmul(phi[:, :, :na].transpose(0, 2, 1), psi[:, :na].conj())
    Ob = numpy.matmul(phi[:, :, na:].transpose(0, 2, 1), psi[:, na:].conj())
    tha = numpy.linalg.solve(Oa, phi[:, :, :na].transpose(0, 2, 1))
    thb = numpy.linalg.solve(Ob, phi[:, :, na:].transpose(0, 2, 1))
    sa, la = _slogdet_prod(Oa)
    sb, lb = _slogdet_prod(Ob)
    det = sa * sb * numpy.exp(la + lb)
    return tha, thb, det


def full_greens_function(ham, tha, thb):
    """G[s] = conj(psi_s) Theta_s (single_det.py:312,319)."""
    na = ham.nup
    Ga = numpy.matmul(ham.psi[:, :na].conj(), tha)
    Gb = numpy.matmul(ham.psi[:, na:].conj(), thb)
    return Ga, Gb


def calc_overlap(ham, phi):
    """walkers/single_det.py:170-199 (nup == ndown assumed there)."""
    na = ham.nup
    Oa = numpy.matmul(ham.psi[:, :na].conj().T, phi[:, :, :na])
    Ob = numpy.matmul(ham.psi[:, na:].conj().T, phi[:, :, na:])
    sa, la = _slogdet_prod(Oa)
    sb, lb = _slogdet_prod(Ob)
    return sa * sb * numpy.exp(la + lb)


def kinetic_real(ham, phi):
    """propagation/operations.py:29-52."""
    na = ham.nup
    out = numpy.empty_like(phi)
    out[:, :, :na] = numpy.matmul(ham.BH1[0], phi[:, :, :na])
    out[:, :, na:] = numpy.matmul(ham.BH1[1], phi[:, :, na:])
    return out


def force_bias(ham, tha, thb):
    """propagation/generic.py:130-152.  Returns (xbar [W,N], vbias [W,N])."""
    W = tha.shape[0]
    M, na, nb = ham.nbasis, ham.nup, ham.ndown
    ra = ham.rchol[:na * M]
    rb = ham.rchol[na * M:]
    vb = numpy.dot(tha.reshape(W, -1), ra) + numpy.dot(thb.reshape(W, -1), rb)
    return -ham.sqrt_dt * (1j * vb - ham.mf_shift), vb


def shift_fields(ham, xi, xbar):
    """propagation/continuous.py:133-158.  xi [W,N] real, xbar [W,N] complex.
    Returns (x [W,N], cmf [W], cfb [W], ntrig)."""
    xbar = numpy.array(xbar, copy=True)
    a = numpy.absolute(xbar)
    big = a > 1.0
    ntrig = int(big.sum())
    xbar[big] = xbar[big] / a[big]
    x = xi - xbar
    cmf = -ham.sqrt_dt * numpy.dot(x, ham.mf_shift)
    cfb = numpy.sum(xi * xbar, axis=1) - 0.5 * numpy.sum(xbar * xbar, axis=1)
    return x, cmf, cfb, ntrig


def construct_vhs(ham, x):
    """propagation/generic.py:164-179.  x [W,N] -> VHS [W,M,M]."""
    M = ham.nbasis
    v = numpy.dot(x, ham.hs_pot.T.astype(numpy.complex128))
    return ham.isqrt_dt * v.reshape(-1, M, M)


def apply_exponential(phi, vhs, order=6):
    """propagation/continuous.py:82-111 (both spin blocks use the same VHS)."""
    out = numpy.array(phi, copy=True)
    temp = numpy.array(phi, copy=True)
    for n in range(1, order + 1):
        temp = numpy.matmul(vhs, temp) / n
        out += temp
    return out


def update_weight_hybrid(ham, weight, ot_old, ot_new, ehyb_old, cfb, cmf, eshift):
    """propagation/continuous.py:202-214,264-292 for one walker (python
    scalars; cmath principal branches as in the reference).
    Returns (weight, ot, hybrid_energy, triggered)."""
    dt = ham.dt
    ratio = ot_new / ot_old
    eh = -(cmath.log(ratio) + cfb + cmf) / dt
    trig = 0
    if abs(eshift) >= 1e-10:
        er = eshift.real if isinstance(eshift, complex) else eshift
        if eh.real > er + ham.ebound:
            eh = er + ham.ebound + 1j * eh.imag
            trig = 1
        elif eh.real < er - ham.ebound:
            eh = er - ham.ebound + 1j * eh.imag
            trig = 1
    imp = cmath.exp(-dt * (0.5 * (eh + ehyb_old) - eshift))
    magn, _ = cmath.polar(imp)
    if not math.isinf(magn):
        dtheta = (-dt * eh - cfb).imag
        weight = weight * magn * max(0, math.cos(dtheta))
        ot = ot_new
    else:
        # the reference raises NameError here (continuous.py:291); SURVEY
        # Appendix A: treat as weight = 0
        weight = 0.0
        ot = ot_new
    return weight, ot, eh, trig


def update_weight_free(ham, weight, phase, cmf, eshift):
    """propagation/continuous.py:194-198 (free projection): the constant terms go into the
    walker's weight and phase.  Returns (weight, phase)."""
    magn, dtheta = cmath.polar(cmath.exp(cmf + ham.dt * eshift))
    return weight * magn, phase * cmath.exp(1j * dtheta)


def reortho(ham, phi):
    """walkers/single_det.py:215-255.  Returns (phi_new, detR [W], log_det [W])."""
    na = ham.nup
    out = numpy.empty_like(phi)
    W = phi.shape[0]
    logdet = numpy.zeros(W)
    for w in range(W):
        for sl in (slice(0, na), slice(na, ham.ne)):
            Q, R = scipy.linalg.qr(phi[w][:, sl], mode='economic')
            d = numpy.diag(R)
            out[w][:, sl] = numpy.dot(Q, numpy.diag(numpy.sign(d)))
            logdet[w] += numpy.sum(numpy.log(numpy.abs(d)))
    return out, numpy.exp(logdet), logdet


def local_energy(ham, tha, thb):
    """estimators/generic.py:156-221 with G from single_det.py:312,319.
    Returns [W,3] complex (E, E1, E2)."""
    W = tha.shape[0]
    M, na, nb = ham.nbasis, ham.nup, ham.ndown
    Ga, Gb = full_greens_function(ham, tha, thb)
    e1b = numpy.sum(ham.H1[0] * Ga, axis=(1, 2)) + numpy.sum(ham.H1[1] * Gb, axis=(1, 2))
    ra = ham.rchol[:na * M]
    rb = ham.rchol[na * M:]
    Xa = numpy.dot(tha.reshape(W, -1), ra)
    Xb = numpy.dot(thb.reshape(W, -1), rb)
    ecoul = numpy.sum(Xa * Xa, axis=1) + numpy.sum(Xb * Xb, axis=1) + 2 * numpy.sum(Xa * Xb, axis=1)
    ra3 = ra.reshape(na, M, -1)
    rb3 = rb.reshape(nb, M, -1)
    exx = numpy.zeros(W, dtype=numpy.complex128)
    for w in range(W):
        Ta = numpy.einsum('ipx,jp->xij', ra3, tha[w], optimize=True)
        Tb = numpy.einsum('ipx,jp->xij', rb3, thb[w], optimize=True)
        exx[w] = numpy.einsum('xij,xji->', Ta, Ta) + numpy.einsum('xij,xji->', Tb, Tb)
    e2b = 0.5 * (ecoul - exx)
    return numpy.stack([e1b + e2b + ham.ecore, e1b + ham.ecore, e2b], axis=1)


# --------------------------------------------------------------------------
# population control (bit-exact integer/selection logic)
# --------------------------------------------------------------------------
def comb_parents(weights, r, target):
    """walkers/handler.py:271-286: sequential python sum, numpy.cumsum and the
    two-pointer sweep.  weights: float64 [W] already divided by scale."""
    total_weight = sum(weights)
    cprobs = numpy.cumsum(weights)
    comb = [(i + r) * (total_weight / target) for i in range(target)]
    parent_ix = numpy.zeros(len(weights), dtype='i')
    iw = 0
    ic = 0
    while ic < len(comb):
        if comb[ic] < cprobs[iw]:
            parent_ix[iw] += 1
            ic += 1
        else:
            iw += 1
    return parent_ix


def comb_pairs(parent_ix):
    """walkers/handler.py:296-301: position-wise zip(clone, kill)."""
    kill = numpy.where(parent_ix == 0)[0]
    clone = numpy.where(parent_ix > 1)[0]
    return list(zip(clone.tolist(), kill.tolist()))


def pair_branch_plan(abs_weights, rand, min_weight, max_weight):
    """walkers/handler.py:340-386 on one rank.  `rand()` yields the uniform
    draws.  Returns (new_weights [W], pairs [(clone, kill)])."""
    w = numpy.array(abs_weights, dtype=numpy.float64)
    sort = numpy.argsort(w, kind='mergesort')
    ws = w[sort].copy()
    s, e = 0, len(ws) - 1
    pairs = []
    state = numpy.ones(len(ws), dtype=int)
    while s < e:
        if ws[s] < min_weight or ws[e] > max_weight:
            wab = ws[s] + ws[e]
            r = rand()
            if r < ws[e] / wab:
                ws[e] = 0.5 * wab
                ws[s] = 0.0
                pairs.append((int(sort[e]), int(sort[s])))
            else:
                ws[s] = 0.5 * wab
                ws[e] = 0.0
                pairs.append((int(sort[s]), int(sort[e])))
            s += 1
            e -= 1
        else:
            break
    new_w = numpy.empty_like(ws)
    new_w[sort] = ws
    # Reference quirk (walkers/handler.py:388-410): on one rank every Isend and
    # Recv carries tag 0, so buffers are matched in POSTING order -- the k-th
    # cloned walker (ascending index) lands in the k-th killed walker
    # (ascending index), not in its pairing partner.
    clones = sorted(c for c, _ in pairs)
    kills = sorted(k for _, k in pairs)
    return new_w, list(zip(clones, kills))


# --------------------------------------------------------------------------
# the driver loop
# --------------------------------------------------------------------------
EST_NAMES = ['uweight', 'weight', 'enumer', 'edenom', 'eproj', 'e1b', 'e2b',
             'ehyb', 'ovlp', 'time']  # estimators/mixed.py:460-469


def exponentiate_matrix(M, order=6):
    """utils/linalg.py:163-170."""
    T = numpy.copy(M)
    out = numpy.identity(M.shape[0], dtype=M.dtype)
    for n in range(1, order + 1):
        out += T
        T = M.dot(T) / (n + 1)
    return out


def back_propagate(ham, phi, configs, nstblz):
    """propagation/generic.py:253-290 (back_propagate_generic) with the propagator matrix of
    generic.py:180-213: for the stored field configurations in REVERSE order,
    phi <- B(c)^dagger phi with B = BH1 exp6(i sqrt(dt) L.c) BH1, a QR re-orthogonalisation
    (utils/linalg.py:82-105) every nstblz applications.  phi [M, ne] is changed in place."""
    M, na = ham.nbasis, ham.nup
    for i, c in enumerate(configs[::-1]):
        vhs = 1j * ham.sqrt_dt * ham.hs_pot.dot(c).reshape(M, M)
        e = exponentiate_matrix(vhs)
        for s, sl in enumerate((slice(0, na), slice(na, ham.ne))):
            B = ham.BH1[s].dot(e).dot(ham.BH1[s])
            phi[:, sl] = numpy.dot(B.conj().T, phi[:, sl])
        if i != 0 and i % nstblz == 0:
            for sl in (slice(0, na), slice(na, ham.ne)):
                if sl.stop > sl.start:
                    Q, R = scipy.linalg.qr(phi[:, sl], mode='economic')
                    phi[:, sl] = Q.dot(numpy.diag(numpy.sign(numpy.diag(R))))
    return phi


def gab(A, B):
    """estimators/greens_function.py:5-38: B (A^dagger B)^-1 A^dagger."""
    inv_o = scipy.linalg.inv((A.conj().T).dot(B))
    return B.dot(inv_o.dot(A.conj().T))


class OracleAFQMC(object):
    """qmc/afqmc.py:200-255 + walkers/handler.py:225-338 + estimators/mixed.py:
    180-289 on structure-of-arrays walker state (single rank)."""

    def __init__(self, ham, nwalkers, nsteps=10, nblocks=10, nstblz=10,
                 npop_control=1, energy_eval_freq=1, exp_order=6,
                 pop_control='comb', min_weight=0.1, max_weight=4.0,
                 verbose_step0=False, free_projection=False, force_bias=True,
                 nbp=0, nsplit=1, init_walker=False, one_rdm=False, use_log_shift=False,
                 restore_weights=None):
        self.ham = ham
        # walkers/handler.py:45-46,456-475, walkers/walker.py:49-52
        self.use_log_shift = use_log_shift
        self.log_shift = self.detR_shift = self.log_detR_shift = 0.0
        self.shift_counter = 1
        W = nwalkers
        self.W = W
        self.nsteps, self.nblocks, self.nstblz = nsteps, nblocks, nstblz
        self.npop_control = npop_control
        self.energy_eval_freq = energy_eval_freq
        self.exp_order = exp_order
        self.pcont = pop_control
        self.min_weight, self.max_weight = min_weight, max_weight
        self.neqlb = int(2.0 / ham.dt)
        self.phi = numpy.array([ham.psi.copy() for _ in range(W)])
        self.weight = numpy.ones(W)
        self.unscaled_weight = numpy.ones(W)
        self.hybrid_energy = numpy.zeros(W, dtype=numpy.complex128)
        self.phase = numpy.ones(W, dtype=numpy.complex128)
        self.free_projection = free_projection
        # continuous.py:30-33: free projection switches the force bias off
        self.force_bias = force_bias and not free_projection
        self.ot = calc_overlap(ham, self.phi).astype(numpy.complex128)
        self.detR = numpy.ones(W)
        self.log_detR = numpy.zeros(W)
        self.total_weight = float(W)
        self.estimates = numpy.zeros(10, dtype=numpy.complex128)
        self.eshift_vec = numpy.array([0, 0], dtype=numpy.complex128)
        self.eshift = 0
        self.nfb_trig = 0
        self.nhe_trig = 0
        self.rows = []
        self.eloc = numpy.zeros((W, 3), dtype=numpy.complex128)
        self.step = 0
        self.last_parent_ix = numpy.ones(W, dtype='i')
        self.verbose_step0 = verbose_step0
        # mixed one-body density matrix (estimators/mixed.py:76,226-229,279-283)
        self.one_rdm = one_rdm
        self.rdm_acc = numpy.zeros((2, ham.nbasis, ham.nbasis))
        self.rdm_out = []
        # back propagation (estimators/back_propagation.py:55-76, walkers/walker.py:43,55-58)
        self.nbp = nbp
        self.init_walker = init_walker
        self.bp_splits = [(i + 1) * (nbp // nsplit) for i in range(nsplit)] if nbp else []
        self.phi_old = self.phi.copy()
        self.configs = numpy.zeros((W, max(nbp, 1), ham.nchol), dtype=numpy.complex128)
        self.cfg_step = numpy.zeros(W, dtype=int)
        # FieldConfig.weight_fac / cos_fac as running products (walkers/stack.py:51-71,118-121)
        self.restore_weights = restore_weights
        self.bp_ph = numpy.ones(W, dtype=numpy.complex128)
        self.bp_cos = numpy.ones(W)
        self.bp_estimates = numpy.zeros(1 + 2 * ham.nbasis * ham.nbasis, dtype=numpy.complex128)
        self.bp_out = []          # (buff_ix, denominator, one_rdm [2, M, M]) per print
        self.estimator_update(0)
        if verbose_step0:
            self.print_step(0, nsteps=1)

    # -- propagation ------------------------------------------------------
    def propagate(self, xi_active, active):
        """One pass of hot loop 1 (afqmc.py:231-236) for the active walkers.
        xi_active [n_active, N]."""
        ham = self.ham
        idx = numpy.where(active)[0]
        if len(idx):
            phi = self.phi[idx]
            tha, thb, ovlp_old = greens_function(ham, phi)
            phi = kinetic_real(ham, phi)
            if self.force_bias:
                xbar, _ = force_bias(ham, tha, thb)
            else:
                xbar = numpy.zeros(xi_active.shape, dtype=numpy.complex128)  # continuous.py:136-138
            x, cmf, cfb, ntrig = shift_fields(ham, xi_active, xbar)
            self.nfb_trig += ntrig
            if self.nbp and not self.free_projection:
                # continuous.py:288-289 -> FieldConfig.update (walkers/stack.py:52-79); a walker
                # whose importance function is not finite skips the update (continuous.py:291-293)
                for k, iw in enumerate(idx):
                    self.configs[iw, self.cfg_step[iw]] = x[k]
                    self.cfg_step[iw] += 1
            vhs = construct_vhs(ham, x)
            phi = apply_exponential(phi, vhs, self.exp_order)
            phi = kinetic_real(ham, phi)
            ovlp_new = calc_overlap(ham, phi)
            self.phi[idx] = phi
            for k, iw in enumerate(idx):
                if self.free_projection:
                    # continuous.py:175-200 (xbar == 0 here: continuous.py:30-33)
                    self.weight[iw], self.phase[iw] = update_weight_free(
                        ham, float(self.weight[iw]), complex(self.phase[iw]), complex(cmf[k]),
                        self.eshift)
                    self.ot[iw] = ovlp_new[k]
                    continue
                w, ot, eh, trig = update_weight_hybrid(
                    ham, float(self.weight[iw]), complex(ovlp_old[k]), complex(ovlp_new[k]),
                    complex(self.hybrid_energy[iw]), complex(cfb[k]), complex(cmf[k]),
                    self.eshift)
                if self.nbp:
                    # continuous.py:273-289: wfac = (I / |I|, cosine_fac) or (0, 0), stored with the fields
                    imp = cmath.exp(-ham.dt * (0.5 * (eh + complex(self.hybrid_energy[iw])) - self.eshift))
                    magn = abs(imp)
                    if not math.isinf(magn):
                        cf = max(0, math.cos((-ham.dt * eh - complex(cfb[k])).imag))
                        self.bp_ph[iw] *= (imp / magn) if magn > 1e-16 else 0.0
                        self.bp_cos[iw] *= cf if magn > 1e-16 else 0.0
                self.weight[iw] = w
                # calc_overlap reports exp(logdet - log_shift) (single_det.py:192)
                self.ot[iw] = ot * math.exp(-self.log_shift)
                self.hybrid_energy[iw] = eh
                self.nhe_trig += trig
        if self.step > 1:
            cap = numpy.abs(self.weight) > self.total_weight * 0.10
            self.weight[cap] = self.total_weight * 0.10

    # -- population control -------------------------------------------------
    def pop_control(self, rand):
        """walkers/handler.py:225-254; `rand()` returns numpy.random.random()."""
        if self.W == 1:
            return
        if self.use_log_shift:
            # Walkers.update_log_ovlp (walkers/handler.py:456-475)
            n, nm1 = self.shift_counter, self.shift_counter - 1
            self.log_shift = (self.log_shift * nm1 + math.log(sum(abs(o) for o in self.ot) / self.W)) / n
            self.detR_shift = (self.detR_shift * nm1 + math.log(sum(abs(x) for x in self.detR) / self.W)) / n
            self.log_detR_shift = (self.log_detR_shift * nm1 + sum(abs(x) for x in self.log_detR) / self.W) / n
            self.shift_counter += 1
        weights = numpy.abs(self.weight)
        total_weight = sum(weights)
        scale = total_weight / self.W
        if total_weight < 1e-8:
            raise RuntimeError("total weight %g < 1e-8" % total_weight)
        self.total_weight = total_weight
        self.unscaled_weight = self.weight.copy()
        self.weight = self.weight / scale
        if self.pcont == 'comb':
            gw = weights / scale
            parent_ix = comb_parents(gw, rand(), self.W)
            self.last_parent_ix = parent_ix
            self._copy_walkers(comb_pairs(parent_ix))
            self.weight[:] = 1.0
        else:
            new_w, pairs = pair_branch_plan(numpy.abs(self.weight), rand,
                                            self.min_weight, self.max_weight)
            for c, k in pairs:
                self.weight[c] = new_w[c]
            self._copy_walkers(pairs)

    def _copy_walkers(self, pairs):
        # full-buffer copy (walkers/walker.py:63-131): every per-walker field
        for c, k in pairs:
            self.phi[k] = self.phi[c]
            self.weight[k] = self.weight[c]
            self.unscaled_weight[k] = self.unscaled_weight[c]
            self.ot[k] = self.ot[c]
            self.hybrid_energy[k] = self.hybrid_energy[c]
            self.phase[k] = self.phase[c]
            self.detR[k] = self.detR[c]
            self.log_detR[k] = self.log_detR[c]
            self.eloc[k] = self.eloc[c]
            self.phi_old[k] = self.phi_old[c]
            self.configs[k] = self.configs[c]
            self.cfg_step[k] = self.cfg_step[c]
            self.bp_ph[k] = self.bp_ph[c]
            self.bp_cos[k] = self.bp_cos[c]

    # -- estimators -----------------------------------------------------------
    def estimator_update(self, step):
        """estimators/mixed.py:211-225 (phaseless, single det)."""
        es = self.estimates
        if step % self.energy_eval_freq == 0:
            tha, thb, _ = greens_function(self.ham, self.phi)
            self.eloc = local_energy(self.ham, tha, thb)
        if self.one_rdm and not self.free_projection:
            # mixed.py:226-229: weight * Re(walker.G), G from the greens_function call above
            # (fresh every step when energy_eval_freq == 1)
            tha, thb, _ = greens_function(self.ham, self.phi)
            Ga, Gb = full_greens_function(self.ham, tha, thb)
            self.rdm_acc[0] += numpy.einsum('w,wpq->pq', self.weight, Ga.real)
            self.rdm_acc[1] += numpy.einsum('w,wpq->pq', self.weight, Gb.real)
        for iw in range(self.W):
            w = self.weight[iw]
            if self.free_projection:
                # estimators/mixed.py:151-177
                wfac = w * self.ot[iw] * self.phase[iw]
                if step % self.energy_eval_freq == 0:
                    E, T, V = self.eloc[iw]
                    es[2] += wfac * E
                    es[5] += wfac * T
                    es[6] += wfac * V
                    es[3] += wfac
                es[0] += self.unscaled_weight[iw]
                es[1] += wfac
                es[7] += wfac * self.hybrid_energy[iw]
                es[8] += w * abs(self.ot[iw])
                continue
            if step % self.energy_eval_freq == 0:
                E, T, V = self.eloc[iw]
                es[2] += w * E.real
                es[5] += w * T.real
                es[6] += w * V.real
                es[3] += w
            es[0] += self.unscaled_weight[iw]
            es[1] += w
            es[8] += w * abs(self.ot[iw])
            es[7] += w * self.hybrid_energy[iw]

    def print_step(self, step, nsteps=None):
        """estimators/mixed.py:252-289."""
        if step % self.nsteps != 0:
            return
        if nsteps is None:
            nsteps = self.nsteps
        gs = self.estimates.copy()
        gs[9] = 0.0
        gs[0:2] /= nsteps
        gs[7:10] /= nsteps
        gs[4] = gs[2]
        gs[4:7] = gs[4:7] / gs[3]
        gs[7] /= gs[1]
        gs[8] /= gs[1]
        self.eshift_vec = numpy.array([gs[7], gs[4]])
        self.rows.append(numpy.concatenate([[step], gs]))
        if self.one_rdm:
            self.rdm_out.append(self.rdm_acc / nsteps / gs[1])     # mixed.py:279-283
            self.rdm_acc[:] = 0
        self.estimates[:] = 0

    # -- back propagation -----------------------------------------------------
    def bp_update(self):
        """estimators/back_propagation.py:127-225 (update_uhf, BP-PhL weights, one_rdm only)
        followed by print_step (:282-333): returns nothing, appends to bp_out."""
        ham = self.ham
        M, na = ham.nbasis, ham.nup
        buff_ix = int(self.cfg_step[0])
        if buff_ix not in self.bp_splits:
            return
        init = self.phi0 if (self.init_walker and hasattr(self, 'phi0')) else ham.psi
        for iw in range(self.W):
            phi_bp = init.copy()
            back_propagate(ham, phi_bp, self.configs[iw, :self.cfg_step[iw]], self.nstblz)
            G = numpy.zeros((2, M, M), dtype=numpy.complex128)
            G[0] = gab(phi_bp[:, :na], self.phi_old[iw][:, :na]).T
            if ham.ne > na:
                G[1] = gab(phi_bp[:, na:], self.phi_old[iw][:, na:]).T
            w = self.weight[iw]
            if self.restore_weights == 'full':          # back_propagation.py:187-196
                w = w * (self.bp_ph[iw] / self.bp_cos[iw])
            elif self.restore_weights is not None:
                w = w * self.bp_ph[iw]
            self.bp_estimates[0] += w
            self.bp_estimates[1:] += w * G.flatten()
            if buff_ix == self.bp_splits[-1]:
                # FieldConfig.reset (walkers/stack.py:122-125), nprop_tot == nbp
                if self.cfg_step[iw] % self.nbp == 0:
                    self.cfg_step[iw] = 0
                    self.bp_ph[iw] = 1.0
                    self.bp_cos[iw] = 1.0
        if buff_ix == self.bp_splits[-1]:
            self.phi_old = self.phi.copy()          # copy_historic_wfn (handler.py:200-203)
        self.bp_out.append((buff_ix, self.bp_estimates[0],
                            self.bp_estimates[1:].reshape(2, M, M).copy()))
        self.bp_estimates[:] = 0

    # -- one driver step ------------------------------------------------------
    def do_step(self, xi_active, rand):
        """Loop body of afqmc.py:223-255.  xi_active: [n_active, N] normals for
        the walkers with |weight| > 1e-8 in list order; rand(): uniform draws."""
        self.step += 1
        step = self.step
        if step % self.nstblz == 0:
            self.phi, detR, logdet = reortho(self.ham, self.phi)
            detR = numpy.exp(logdet - self.detR_shift)      # single_det.py:250
            self.detR = detR
            self.log_detR = self.log_detR + numpy.log(detR)
            self.ot = self.ot / detR
            if self.free_projection:
                # walkers/handler.py:178-181: polar(detR), detR real positive
                self.weight = self.weight * numpy.abs(detR)
                self.phase = self.phase * numpy.exp(1j * numpy.angle(detR))
        active = numpy.abs(self.weight) > 1e-8
        self.propagate(xi_active, active)
        if step % self.npop_control == 0:
            self.pop_control(rand)
        self.estimator_update(step)
        if self.nbp:
            self.bp_update()
        self.print_step(step)
        if step < self.neqlb:
            # afqmc.py:251-252 get_shift(propagators.hybrid): hybrid energy, or the projected
            # energy with the local-energy weight update (mixed.py:345-360)
            self.eshift = self.eshift_vec[0 if getattr(self, 'hybrid', True) else 1].real
        else:
            self.eshift += (self.eshift_vec[0].real - self.eshift)

    def active_mask(self):
        return numpy.abs(self.weight) > 1e-8

    def run(self, seed, nsteps_total=None):
        """Whole run drawing from the GLOBAL legacy numpy stream as the
        reference does: seed, then per step one batched
        normal(size=(n_active, N)) and one random() per comb."""
        numpy.random.seed(seed)
        total = self.nsteps * self.nblocks if nsteps_total is None else nsteps_total
        N = self.ham.nchol
        for _ in range(total):
            n_active = int(self.active_mask().sum())
            xi = numpy.random.normal(0.0, 1.0, (n_active, N))
            self.do_step(xi, numpy.random.random)
        return numpy.array(self.rows)

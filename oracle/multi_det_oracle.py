"""TEST INFRASTRUCTURE -- CPU (numpy) restatement of the reference hot path for
MULTI-DETERMINANT trials and of the local-energy weight
update (row A8'), which the reference only supports with MultiDetWalker.

Restated from (paths under /root/reference/pauxy/):
  walkers/multi_det.py:27-300          MultiDetWalker: per-determinant overlaps, weights and Green's
                                       functions, overlap = sum_i conj(c_i) det(O_i_up) det(O_i_dn)
  propagation/generic.py:38-43,82-86   mean-field shift through trial.contract_one_body
  propagation/generic.py:154-157       force bias through walker.contract_one_body
  propagation/continuous.py:216-231,294-318   local-energy bound and weight update
  estimators/mixed.py:439-448          local_energy_multi_det
  estimators/generic.py:398-434        local_energy_generic_cholesky (full G form; evaluated here in
                                       the equivalent half-rotated form per determinant)
  trial_wavefunction/multi_slater.py:153-176,190-205,235-259   trial energy, PHMSD orbitals,
                                       contract_one_body
  estimators/ci.py:187-300             Slater-Condon matrix elements for orthogonal expansions

Only tests/ and the golden generator import this module.  PARITY PINNED by
tests/test_oracle_cpu.py against traces of the unmodified reference
(tests/golden/md_*.npz) and the reference's own known answers
propagation/tests/test_generic.py:70,92.
"""
import cmath
import math

import numpy
import scipy.linalg

from oracle import afqmc_oracle as orc


def phmsd_orbitals(nbasis, nup, ndown, occa, occb):
    """trial_wavefunction/multi_slater.py:190-205: determinants of an orthogonal (particle-hole)
    expansion as columns of the identity."""
    D = len(occa)
    psi = numpy.zeros((D, nbasis, nup + ndown), dtype=numpy.complex128)
    I = numpy.eye(nbasis, dtype=numpy.complex128)
    for i in range(D):
        psi[i, :, :nup] = I[:, list(occa[i])]
        psi[i, :, nup:] = I[:, list(occb[i])]
    return psi


def _gab_mod_ovlp(A, B):
    """estimators/greens_function.py gab_mod_ovlp: O = B^T conj(A), GH = O^-1 B^T, G = conj(A) GH."""
    inv_O = scipy.linalg.inv(numpy.dot(B.T, A.conj()))
    GH = numpy.dot(inv_O, B.T)
    return numpy.dot(A.conj(), GH), GH, inv_O


class MultiDetHamiltonian(orc.Hamiltonian):
    """Hamiltonian + multi-determinant trial: per-determinant half-rotated Cholesky vectors and
    one-body integrals, mean-field shift from the trial's one-body expectation values."""

    def __init__(self, h1e, hs_pot, ecore, nelec, dt, coeffs, psi, init=None, ortho_occs=None):
        self.coeffs = numpy.array(coeffs, dtype=numpy.complex128)
        self.psis = numpy.array(psi, dtype=numpy.complex128)
        self.ndets = len(self.coeffs)
        self.ortho_occs = ortho_occs      # (occa, occb) for an orthogonal expansion, else None
        self._init = init
        orc.Hamiltonian.__init__(self, h1e, hs_pot, ecore, nelec, dt, psi=self.psis[0])

    # Hamiltonian.__init__ computes single-determinant quantities from psi[0]; the multi-det ones
    # replace them here (called lazily so that the base constructor has run)
    def setup_multi_det(self):
        M, na, nb = self.nbasis, self.nup, self.ndown
        chol3 = self.hs_pot.reshape((M, M, -1))
        self.rchols, self.h1rots = [], []
        for psi in self.psis:
            rup = numpy.tensordot(psi[:, :na].conj(), chol3, axes=((0), (0))).reshape((na * M, -1))
            rdn = numpy.tensordot(psi[:, na:].conj(), chol3, axes=((0), (0))).reshape((nb * M, -1))
            self.rchols.append(numpy.concatenate([rup, rdn]))
            self.h1rots.append(numpy.concatenate([numpy.dot(psi[:, :na].conj().T, self.H1[0]),
                                                  numpy.dot(psi[:, na:].conj().T, self.H1[1])]))
        self.init = self.psis[0].copy() if self._init is None else numpy.array(self._init)
        # propagation/generic.py:82-86
        self.mf_shift = 1j * numpy.array([self.contract_one_body(self.hs_pot[:, n].reshape(M, M))
                                          for n in range(self.nchol)])
        shift = 1j * self.hs_pot.dot(self.mf_shift).reshape(M, M)
        H1 = self.h1e_mod - numpy.array([shift, shift])
        self.BH1 = numpy.array([scipy.linalg.expm(-0.5 * self.dt * H1[0]),
                                scipy.linalg.expm(-0.5 * self.dt * H1[1])])
        self.mf_core = self.ecore + 0.5 * numpy.dot(self.mf_shift, self.mf_shift)
        return self

    # ---- trial-only expectation values (setup) --------------------------------
    def _spin_occs(self):
        occa, occb = self.ortho_occs
        return [numpy.sort(list(a) + [i + self.nbasis for i in c]) for a, c in zip(occa, occb)]

    def contract_one_body(self, ints):
        """trial_wavefunction/multi_slater.py:235-259."""
        numer, denom = 0.0, 0.0
        na = self.nup
        occs = self._spin_occs() if self.ortho_occs is not None else None
        for i in range(self.ndets):
            for j in range(self.ndets):
                cfac = self.coeffs[i].conj() * self.coeffs[j].conj()      # sic (reference)
                if occs is not None:
                    numer += cfac * one_body_matel(ints, occs[i], occs[j])
                    if i == j:
                        denom += self.coeffs[i].conj() * self.coeffs[i].conj()
                else:
                    di, dj = self.psis[i], self.psis[j]
                    ga, _, ioa = _gab_mod_ovlp(di[:, :na], dj[:, :na])
                    gb, _, iob = _gab_mod_ovlp(di[:, na:], dj[:, na:])
                    ovlp = 1.0 / (scipy.linalg.det(ioa) * scipy.linalg.det(iob))
                    numer += cfac * ovlp * numpy.dot(ints.ravel(), ga.ravel() + gb.ravel())
                    denom += cfac * ovlp
        return numer / denom

    def hijkl(self, i, j, k, l):
        """systems/generic.py:168-171."""
        M = self.nbasis
        return numpy.dot(self.hs_pot[i * M + k, :], self.hs_pot[j * M + l, :])

    def trial_energy(self):
        """multi_slater.py:153-176: variational energy of the trial (orthogonal expansion: Slater-
        Condon rules, estimators/mixed.py:537-572; otherwise the double sum over determinant pairs,
        mixed.py:511-535)."""
        if self.ortho_occs is not None:
            occs = self._spin_occs()
            ev, denom = numpy.zeros(3, dtype=numpy.complex128), 0.0
            for i in range(self.ndets):
                denom += self.coeffs[i].conj() * self.coeffs[i]
                for j in range(i + 1):
                    e = self.coeffs[i].conj() * self.coeffs[j] * hmatel(self, occs[i], occs[j])
                    ev += e
                    if j < i:
                        ev += e
            return ev / denom
        na = self.nup
        energies, denom = 0.0, 0.0
        for Bi, ci in zip(self.psis, self.coeffs):
            for Aj, cj in zip(self.psis, self.coeffs):
                _, ghu, iou = _gab_mod_ovlp(Bi[:, :na], Aj[:, :na])
                _, ghd, iod = _gab_mod_ovlp(Bi[:, na:], Aj[:, na:])
                ovlp = 1.0 / (scipy.linalg.det(iou) * scipy.linalg.det(iod))
                w = ci.conj() * cj * ovlp
                ham1 = _SingleDetView(self, Bi)
                e = orc.local_energy(ham1, ghu[None], ghd[None])[0]
                energies = energies + w * e
                denom += w
        return energies / denom


class _SingleDetView(object):
    """The single-determinant members orc.local_energy reads, for one determinant."""

    def __init__(self, ham, psi):
        M, na, nb = ham.nbasis, ham.nup, ham.ndown
        self.nbasis, self.nup, self.ndown, self.ecore, self.H1, self.psi = M, na, nb, ham.ecore, ham.H1, psi
        chol3 = ham.hs_pot.reshape((M, M, -1))
        rup = numpy.tensordot(psi[:, :na].conj(), chol3, axes=((0), (0))).reshape((na * M, -1))
        rdn = numpy.tensordot(psi[:, na:].conj(), chol3, axes=((0), (0))).reshape((nb * M, -1))
        self.rchol = numpy.concatenate([rup, rdn])


# ---- Slater-Condon rules (estimators/ci.py:187-300) --------------------------------
def _map_orb(orb, nbasis):
    return (orb, 0) if orb // nbasis == 0 else (orb - nbasis, 1)


def _perm(from_orb, to_orb, di, dj):
    nmove, perm = 0, 0
    for o in from_orb:
        perm += int(numpy.where(dj == o)[0][0]) - nmove
        nmove += 1
    nmove = 0
    for o in to_orb:
        perm += int(numpy.where(di == o)[0][0]) - nmove
        nmove += 1
    return perm % 2 == 1


def one_body_matel(ints, di, dj):
    """estimators/ci.py get_one_body_matel."""
    from_orb = sorted(set(dj) - set(di))
    to_orb = sorted(set(di) - set(dj))
    nb = ints.shape[-1]
    if len(from_orb) == 0:
        return sum(ints[_map_orb(o, nb)[0], _map_orb(o, nb)[0]] for o in di)
    if len(from_orb) == 1:
        i, si = _map_orb(from_orb[0], nb)
        a, sa = _map_orb(to_orb[0], nb)
        m = ints[i, a] if si == sa else 0.0
        return -m if _perm(from_orb, to_orb, di, dj) else m
    return 0.0


def hmatel(ham, di, dj):
    """estimators/ci.py get_hmatel: (H, one-body, two-body) matrix element between determinants
    given as sorted spin-orbital lists (beta orbitals offset by nbasis)."""
    nb = ham.nbasis
    from_orb = sorted(set(dj) - set(di))
    to_orb = sorted(set(di) - set(dj))
    nex = len(from_orb)
    if nex > 2:
        return numpy.zeros(3)
    perm = _perm(from_orb, to_orb, di, dj)
    if nex == 0:
        e1b, e2b = ham.ecore, 0.0
        for x in range(len(di)):
            ii, si = _map_orb(di[x], nb)
            e1b += ham.H1[0, ii, ii]
            for y in range(x + 1, len(di)):
                jj, sj = _map_orb(di[y], nb)
                e2b += ham.hijkl(ii, jj, ii, jj)
                if si == sj:
                    e2b -= ham.hijkl(ii, jj, jj, ii)
        return numpy.array([e1b + e2b, e1b, e2b])
    if nex == 1:
        ii, si = _map_orb(from_orb[0], nb)
        aa, sa = _map_orb(to_orb[0], nb)
        e1b, e2b = ham.H1[0, ii, aa], 0.0
        for o in di:
            oj, soj = _map_orb(o, nb)
            if 2 * oj + soj != 2 * ii + si:
                e2b += ham.hijkl(ii, oj, aa, oj)
                if soj == si:
                    e2b -= ham.hijkl(ii, oj, oj, aa)
        v = numpy.array([e1b + e2b, e1b, e2b])
        return -v if perm else v
    ii, si = _map_orb(from_orb[0], nb)
    jj, sj = _map_orb(from_orb[1], nb)
    aa, sa = _map_orb(to_orb[0], nb)
    bb, sb = _map_orb(to_orb[1], nb)
    h = 0.0
    if si == sa:
        h = ham.hijkl(ii, jj, aa, bb)
    if si == sb:
        h -= ham.hijkl(ii, jj, bb, aa)
    h = -h if perm else h
    return numpy.array([h, 0.0, h])


# ---- per-walker stages -------------------------------------------------------------
def md_greens(ham, phi):
    """walkers/multi_det.py:198-231 for a batch: per determinant (theta_up [W,na,M], theta_dn
    [W,nb,M]) and the determinant overlaps ovlps [W, D] (without the CI coefficients)."""
    thetas, ovlps = [], []
    for psi in ham.psis:
        view = _PsiOnly(ham, psi)
        tha, thb, det = orc.greens_function(view, phi)
        thetas.append((tha, thb))
        ovlps.append(det)
    return thetas, numpy.array(ovlps).T


class _PsiOnly(object):
    def __init__(self, ham, psi):
        self.nup, self.psi = ham.nup, psi


def md_overlaps(ham, phi):
    """walkers/multi_det.py:141-166 (calc_overlap): ovlps [W, D]."""
    return numpy.array([orc.calc_overlap(_PsiOnly(ham, psi), phi) for psi in ham.psis]).T


def md_force_bias(ham, thetas, ovlps):
    """propagation/generic.py:154-157 + multi_det.py:292-300: vbias_n = sum_i ofac_i tr(V_n (G_i_up +
    G_i_dn)) / sum_i ofac_i, ofac_i = conj(c_i) ovlp_i; tr(V_n G_i) evaluated in the half-rotated
    form.  Returns xbar [W, N]."""
    W = ovlps.shape[0]
    M, na = ham.nbasis, ham.nup
    ofac = ham.coeffs.conj()[None, :] * ovlps
    numer = numpy.zeros((W, ham.nchol), dtype=numpy.complex128)
    for i, (tha, thb) in enumerate(thetas):
        r = ham.rchols[i]
        vb = numpy.dot(tha.reshape(W, -1), r[:na * M]) + numpy.dot(thb.reshape(W, -1), r[na * M:])
        numer += ofac[:, i:i + 1] * vb
    vbias = numer / ofac.sum(axis=1)[:, None]
    return -ham.sqrt_dt * (1j * vbias - ham.mf_shift)


def md_det_energies(ham, thetas):
    """local energy (E, E1, E2) of every (walker, determinant): [W, D, 3]."""
    out = []
    for i, (tha, thb) in enumerate(thetas):
        view = _SingleDetView(ham, ham.psis[i])
        out.append(orc.local_energy(view, tha, thb))
    return numpy.array(out).transpose(1, 0, 2)


def md_local_energy(ham, det_energies, weights):
    """estimators/mixed.py:439-448: sum_i w_i E_i / sum_i w_i, w_i = conj(c_i) ovlp_i."""
    return numpy.einsum('wd,wdk->wk', weights, det_energies) / weights.sum(axis=1)[:, None]


def update_weight_local_energy(ham, weight, ot_old, ot_new, eloc, eloc_old, eshift):
    """propagation/continuous.py:216-231,294-318 for one walker.  Returns (weight, ot, eloc,
    triggered); eloc (complex, un-bounded) becomes walker.eloc."""
    dt = ham.dt
    ratio = ot_new / ot_old
    trig = 0
    re_eloc = eloc
    if abs(eshift) >= 1e-10:
        er = complex(eshift).real
        if eloc.real > er + ham.ebound:
            re_eloc = er + ham.ebound
            trig = 1
        elif eloc.real < er - ham.ebound:
            re_eloc = er - ham.ebound
            trig = 1
    magn = numpy.exp(-0.5 * dt * (re_eloc + eloc_old - eshift).real)
    if not math.isinf(magn):
        dtheta = cmath.phase(ratio)
        weight = weight * magn * max(0, math.cos(dtheta))
    else:
        weight = 0.0
    return weight, ot_new, eloc, trig


class OracleMultiDet(orc.OracleAFQMC):
    """Driver loop of orc.OracleAFQMC with MultiDetWalker semantics; hybrid=False selects the
    local-energy weight update."""

    def __init__(self, ham, nwalkers, hybrid=True, **kw):
        self.hybrid = hybrid
        self.ndets = ham.ndets
        orc.OracleAFQMC.__init__(self, ham, nwalkers, **kw)

    def reset_walkers(self):
        """MultiDetWalker.__init__: phi = trial.init, ot = overlap_direct."""
        ham = self.ham
        self.phi = numpy.array([ham.init.copy() for _ in range(self.W)])
        ov = md_overlaps(ham, self.phi)
        self.ot = (ham.coeffs.conj()[None, :] * ov).sum(axis=1)
        self.walker_eloc = numpy.zeros(self.W, dtype=numpy.complex128)   # walker.eloc = 0 (walker.py:37)
        self.estimates[:] = 0
        self.estimator_update(0)

    def propagate(self, xi_active, active):
        ham = self.ham
        idx = numpy.where(active)[0]
        if len(idx):
            phi = self.phi[idx]
            thetas, ovlps = md_greens(ham, phi)
            ovlp_old = (ham.coeffs.conj()[None, :] * ovlps).sum(axis=1)
            if not self.hybrid:
                e_old = md_det_energies(ham, thetas)      # E(G_i) of the walker BEFORE the step
            phi = orc.kinetic_real(ham, phi)
            xbar = md_force_bias(ham, thetas, ovlps)
            x, cmf, cfb, ntrig = orc.shift_fields(ham, xi_active, xbar)
            self.nfb_trig += ntrig
            vhs = orc.construct_vhs(ham, x)
            phi = orc.apply_exponential(phi, vhs, self.exp_order)
            phi = orc.kinetic_real(ham, phi)
            ov_new = md_overlaps(ham, phi)
            w_new = ham.coeffs.conj()[None, :] * ov_new
            ovlp_new = w_new.sum(axis=1)
            self.phi[idx] = phi
            if not self.hybrid:
                # continuous.py:296: walker.local_energy uses the Green's functions left by
                # greens_function at the top of the step with the weights calc_overlap just updated
                eloc = md_local_energy(ham, e_old, w_new)[:, 0]
            for k, iw in enumerate(idx):
                if self.hybrid:
                    w, ot, eh, trig = orc.update_weight_hybrid(
                        ham, float(self.weight[iw]), complex(ovlp_old[k]), complex(ovlp_new[k]),
                        complex(self.hybrid_energy[iw]), complex(cfb[k]), complex(cmf[k]), self.eshift)
                    self.hybrid_energy[iw] = eh
                else:
                    w, ot, el, trig = update_weight_local_energy(
                        ham, float(self.weight[iw]), complex(ovlp_old[k]), complex(ovlp_new[k]),
                        complex(eloc[k]), complex(self.walker_eloc[iw]), self.eshift)
                    self.walker_eloc[iw] = el
                self.weight[iw] = w
                self.ot[iw] = ot
                self.nhe_trig += trig
        if self.step > 1:
            cap = numpy.abs(self.weight) > self.total_weight * 0.10
            self.weight[cap] = self.total_weight * 0.10

    def _copy_walkers(self, pairs):
        for c, k in pairs:
            self.walker_eloc[k] = self.walker_eloc[c]
        orc.OracleAFQMC._copy_walkers(self, pairs)

    def estimator_update(self, step):
        """estimators/mixed.py:211-225 with MultiDetWalker.greens_function / local_energy."""
        es = self.estimates
        if not hasattr(self, 'walker_eloc'):
            return            # base constructor: the multi-det state is set up by reset_walkers
        if step % self.energy_eval_freq == 0:
            thetas, ovlps = md_greens(self.ham, self.phi)
            w = self.ham.coeffs.conj()[None, :] * ovlps
            self.eloc = md_local_energy(self.ham, md_det_energies(self.ham, thetas), w)
        for iw in range(self.W):
            w = self.weight[iw]
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

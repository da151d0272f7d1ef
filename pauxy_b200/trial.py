"""Host-side trial wavefunction (setup-only data provider), one or several determinants.

Mirrors the members of pauxy.trial_wavefunction.multi_slater.MultiSlater the
hot path reads (multi_slater.py:17-98, half_rotate :267-420, rot_hs_pot
:436-448) for ndets == 1, and the RHF guess of
pauxy/trial_wavefunction/utils.py:64-77.
"""
import numpy
import scipy.linalg


def gab_mod(A, B):
    """pauxy/estimators/greens_function.py:41-67."""
    O = numpy.dot(B.T, A.conj())
    GHalf = numpy.dot(scipy.linalg.inv(O), B.T)
    G = numpy.dot(A.conj(), GHalf)
    return (G, GHalf)


class MultiSlater(object):
    """Single- or multi-determinant trial (multi_slater.py:17-98).

    wfn = (coeffs, psi[ndets, M, ne])  non-orthogonal expansion, or one determinant psi[M, ne];
    wfn = (coeffs, occa, occb)         particle-hole (orthogonal) expansion: occupied orbital lists
                                       per determinant (multi_slater.py:190-205).
    Orbitals and CI coefficients may be complex (complex orbitals select PXB_FLAG_COMPLEX_CHOLESKY)."""

    def __init__(self, system, wfn, init=None, options=None, verbose=False):
        self.name = "MultiSlater"
        self.type = "MultiSlater"
        self.verbose = verbose
        na, nb = system.nup, system.ndown
        M = system.nbasis
        self.ortho_expansion = len(wfn) == 3
        if self.ortho_expansion:
            coeffs, occa, occb = wfn
            self.occa = [tuple(int(x) for x in o) for o in occa]
            self.occb = [tuple(int(x) for x in o) for o in occb]
            psi = numpy.zeros((len(coeffs), M, na + nb), dtype=numpy.complex128)
            I = numpy.eye(M, dtype=numpy.complex128)
            for i, (oa, ob) in enumerate(zip(self.occa, self.occb)):
                psi[i, :, :na] = I[:, list(oa)]
                psi[i, :, na:] = I[:, list(ob)]
            # alpha orbitals first, beta offset by nbasis, sorted (multi_slater.py:198-199)
            self.spin_occs = [numpy.sort(list(oa) + [p + M for p in ob])
                              for oa, ob in zip(self.occa, self.occb)]
        else:
            coeffs, psi = wfn
            psi = numpy.asarray(psi)
        self.coeffs = numpy.array(coeffs, dtype=numpy.complex128)
        if (options or {}).get('split_trial_local_energy', False):
            raise NotImplementedError("pauxy_b200: split_trial_local_energy is not built")
        self.split_trial_local_energy = False
        if psi.ndim == 3 and psi.shape[0] == 1:
            psi = psi[0]
        self.ndets = 1 if psi.ndim == 2 else psi.shape[0]
        if len(self.coeffs) != self.ndets:
            raise ValueError("MultiSlater: %d coefficients for %d determinants" % (len(self.coeffs), self.ndets))
        # Walkers.__init__ strips the determinant axis of a single determinant (handler.py:57-61)
        self.psi = numpy.array(psi, dtype=numpy.complex128)
        if self.ndets == 1:
            Ga, Gha = gab_mod(self.psi[:, :na], self.psi[:, :na])
            Gb, Ghb = gab_mod(self.psi[:, na:], self.psi[:, na:])
            self.G = numpy.array([Ga, Gb])
            self.GH = [Gha, Ghb]
        else:
            self.G = None           # multi_slater.py:64-66
            self.GH = None
        first = self.psi if self.ndets == 1 else self.psi[0]
        self.init = numpy.array(init if init is not None else first, dtype=numpy.complex128)
        self._nalpha, self._nbeta = na, nb
        self._nbasis = M
        self._rchol = None
        self._rot_hs_pot = None
        self._eri = None
        self._UVT = None
        self.energy = None

    def det(self, idet=0):
        return self.psi if self.ndets == 1 else self.psi[idet]

    def half_rotate(self, system, comm=None):
        """R[(i,p),n] = sum_m conj(psi[m,i]) L[(m,p),n]; spin-up rows first
        (multi_slater.py:402-409), one block per determinant.  Stored complex128 as the reference
        does; `_rchol` is [ne*M, N] for one determinant and [ndets, ne*M, N] for several."""
        M, na, nb = system.nbasis, system.nup, system.ndown
        chol = system.chol_vecs.reshape((M, M, -1))
        blocks = []
        for i in range(self.ndets):
            psi = self.det(i)
            rup = numpy.tensordot(psi[:, :na].conj(), chol, axes=((0), (0))).reshape((na * M, -1))
            rdn = numpy.tensordot(psi[:, na:].conj(), chol, axes=((0), (0))).reshape((nb * M, -1))
            blocks.append(numpy.concatenate([rup, rdn]).astype(numpy.complex128))
        self._rchol = blocks[0] if self.ndets == 1 else numpy.array(blocks)
        self._rot_hs_pot = self._rchol

    def rchol(self, idet=0):
        return self._rchol if self.ndets == 1 else self._rchol[idet]

    def rot_hs_pot(self, idet=0, spin=None):
        alpha = self._nbasis * self._nalpha
        r = self.rchol(idet)
        if spin is None:
            return r
        return r[:alpha] if spin == 0 else r[alpha:]

    def half_rotated_h1(self, system, idet=0):
        """h1rot[s] = psi_s^dagger H1[s] stacked (up rows, then down):
        sum(h1rot * Theta) == sum(H1[s] * G[s]) of estimators/generic.py:178."""
        na = system.nup
        psi = self.det(idet)
        up = numpy.dot(psi[:, :na].conj().T, system.H1[0])
        dn = numpy.dot(psi[:, na:].conj().T, system.H1[1])
        return numpy.concatenate([up, dn]).astype(numpy.complex128)

    def _pair(self, i, j):
        """(overlap, G_up + G_dn, half-rotated Green's functions) of the determinant pair (i, j)
        (estimators/greens_function.py gab_mod_ovlp)."""
        na = self._nalpha
        di, dj = self.det(i), self.det(j)
        out, ovlp = [], 1.0
        for sl in (slice(0, na), slice(na, None)):
            O = numpy.dot(dj[:, sl].T, di[:, sl].conj())
            ovlp = ovlp * scipy.linalg.det(O)
            gh = numpy.dot(scipy.linalg.inv(O), dj[:, sl].T)
            out.append((numpy.dot(di[:, sl].conj(), gh), gh))
        return ovlp, out

    def contract_one_body(self, ints):
        """<psi_T| sum ints[p,q] a_p^+ a_q |psi_T> / <psi_T|psi_T> as the reference evaluates it for
        the mean-field shift (multi_slater.py:235-259; both coefficients conjugated there)."""
        from .ci import one_body_element
        numer, denom = 0.0, 0.0
        for i in range(self.ndets):
            for j in range(self.ndets):
                cfac = self.coeffs[i].conj() * self.coeffs[j].conj()
                if self.ortho_expansion:
                    numer += cfac * one_body_element(ints, self.spin_occs[i], self.spin_occs[j])
                    if i == j:
                        denom += cfac
                else:
                    ovlp, g = self._pair(i, j)
                    numer += cfac * ovlp * numpy.dot(ints.ravel(), (g[0][0] + g[1][0]).ravel())
                    denom += cfac * ovlp
        return numer / denom

    def calculate_energy(self, system):
        """Variational energy of the trial (multi_slater.py:153-176): one determinant from its own
        Green's function (estimators/generic.py:156-221); an orthogonal expansion by the
        Slater-Condon rules (estimators/mixed.py:537-572); a non-orthogonal one as the double sum
        over determinant pairs (mixed.py:511-535).  Host numpy."""
        M, na, nb = system.nbasis, system.nup, system.ndown
        if self._rchol is None:
            self.half_rotate(system)

        def det_energy(rchol, gha, ghb, h1rot):
            e1b = numpy.sum(h1rot[:na] * gha) + numpy.sum(h1rot[na:] * ghb)
            ra, rb = rchol[:na * M], rchol[na * M:]
            Xa, Xb = ra.T.dot(gha.ravel()), rb.T.dot(ghb.ravel())
            ecoul = numpy.dot(Xa, Xa) + numpy.dot(Xb, Xb) + 2 * numpy.dot(Xa, Xb)
            Ta = numpy.einsum('ipx,jp->xij', ra.reshape(na, M, -1), gha, optimize=True)
            Tb = numpy.einsum('ipx,jp->xij', rb.reshape(nb, M, -1), ghb, optimize=True)
            exx = numpy.einsum('xij,xji->', Ta, Ta) + numpy.einsum('xij,xji->', Tb, Tb)
            e2b = 0.5 * (ecoul - exx)
            return numpy.array([e1b + e2b + system.ecore, e1b + system.ecore, e2b])
        if self.ndets == 1:
            e = det_energy(self._rchol, self.GH[0], self.GH[1], self.half_rotated_h1(system))
        elif self.ortho_expansion:
            from .ci import hamiltonian_element
            e = numpy.zeros(3, dtype=numpy.complex128)
            denom = 0.0
            for i in range(self.ndets):
                denom += self.coeffs[i].conj() * self.coeffs[i]
                for j in range(i + 1):
                    hij = self.coeffs[i].conj() * self.coeffs[j] * hamiltonian_element(
                        system, self.spin_occs[i], self.spin_occs[j])
                    e += hij if j == i else 2 * hij     # "use Hermiticity" (mixed.py:567-571)
            e = e / denom
        else:
            e, denom = 0.0, 0.0
            for i in range(self.ndets):
                for j in range(self.ndets):
                    ovlp, g = self._pair(i, j)
                    w = self.coeffs[i].conj() * self.coeffs[j] * ovlp
                    e = e + w * det_energy(self.rchol(i), g[0][1], g[1][1],
                                           self.half_rotated_h1(system, i))
                    denom += w
            e = e / denom
        self.energy, self.e1b, self.e2b = e[0], e[1], e[2]
        return self.energy


def get_trial_wavefunction(system, options=None, comm=None, scomm=None, verbose=False):
    """pauxy/trial_wavefunction/utils.py:9-77 for name == 'MultiSlater' without a
    wavefunction file: RHF guess, identity columns."""
    options = options or {}
    name = options.get('name', 'MultiSlater')
    if name != 'MultiSlater':
        raise NotImplementedError("pauxy_b200: only the MultiSlater single-determinant trial "
                                  "is on the hot path (got %r)" % name)
    if options.get('filename') is not None:
        # trial_wavefunction/utils.py:26-58: QMCPACK-format wavefunction file (needs h5py)
        from . import io
        wfn, psi0 = io.read_qmcpack_wfn(options['filename'], nelec=system.nelec)
        ndets = options.get('ndets', None)
        if ndets is not None:
            wfn = tuple(w[:ndets] for w in wfn)
        trial = MultiSlater(system, wfn, init=psi0, options=options, verbose=verbose)
        trial.half_rotate(system, scomm)
        return trial
    na, nb = system.nup, system.ndown
    wfn = numpy.zeros((1, system.nbasis, na + nb), dtype=numpy.complex128)
    I = numpy.identity(system.nbasis, dtype=numpy.complex128)
    wfn[0, :, :na] = I[:, :na]
    wfn[0, :, na:] = I[:, :nb]
    trial = MultiSlater(system, (numpy.array([1.0 + 0j]), wfn), options=options, verbose=verbose)
    trial.half_rotate(system, scomm)
    return trial

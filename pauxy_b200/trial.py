"""Host-side single-determinant trial wavefunction (setup-only data provider).

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
    def __init__(self, system, wfn, init=None, options=None, verbose=False):
        self.name = "MultiSlater"
        self.type = "MultiSlater"
        self.verbose = verbose
        coeffs, psi = wfn
        self.coeffs = numpy.array(coeffs, dtype=numpy.complex128)
        psi = numpy.asarray(psi)
        if psi.ndim == 3:
            if psi.shape[0] != 1:
                raise NotImplementedError("pauxy_b200: multi-determinant trials are outside the "
                                          "hot path built here (SURVEY.md section 8f)")
            psi = psi[0]
        self.ndets = 1
        na, nb = system.nup, system.ndown
        # Walkers.__init__ strips the determinant axis (pauxy/walkers/handler.py:57-61)
        self.psi = numpy.array(psi, dtype=numpy.complex128)
        Ga, Gha = gab_mod(self.psi[:, :na], self.psi[:, :na])
        Gb, Ghb = gab_mod(self.psi[:, na:], self.psi[:, na:])
        self.G = numpy.array([Ga, Gb])
        self.GH = [Gha, Ghb]
        self.init = numpy.array(init if init is not None else self.psi, dtype=numpy.complex128)
        self._nalpha, self._nbeta = na, nb
        self._nbasis = system.nbasis
        self._rchol = None
        self._rot_hs_pot = None
        self._eri = None
        self._UVT = None
        self.energy = None

    def half_rotate(self, system, comm=None):
        """R[(i,p),n] = sum_m conj(psi[m,i]) L[(m,p),n]; spin-up rows first
        (multi_slater.py:402-409).  Stored complex128 as the reference does."""
        M, na, nb = system.nbasis, system.nup, system.ndown
        chol = system.chol_vecs.reshape((M, M, -1))
        rup = numpy.tensordot(self.psi[:, :na].conj(), chol, axes=((0), (0))).reshape((na * M, -1))
        rdn = numpy.tensordot(self.psi[:, na:].conj(), chol, axes=((0), (0))).reshape((nb * M, -1))
        self._rchol = numpy.concatenate([rup, rdn]).astype(numpy.complex128)
        self._rot_hs_pot = self._rchol

    def rot_hs_pot(self, idet=0, spin=None):
        alpha = self._nbasis * self._nalpha
        if spin is None:
            return self._rot_hs_pot
        return self._rot_hs_pot[:alpha] if spin == 0 else self._rot_hs_pot[alpha:]

    def half_rotated_h1(self, system):
        """h1rot[s] = psi_s^dagger H1[s] stacked (up rows, then down):
        sum(h1rot * Theta) == sum(H1[s] * G[s]) of estimators/generic.py:178."""
        na = system.nup
        up = numpy.dot(self.psi[:, :na].conj().T, system.H1[0])
        dn = numpy.dot(self.psi[:, na:].conj().T, system.H1[1])
        return numpy.concatenate([up, dn]).astype(numpy.complex128)

    def calculate_energy(self, system):
        """Variational energy of the trial from its own Green's function
        (host numpy; same contraction as estimators/generic.py:156-221)."""
        M, na, nb = system.nbasis, system.nup, system.ndown
        e1b = numpy.sum(system.H1[0] * self.G[0]) + numpy.sum(system.H1[1] * self.G[1])
        ra, rb = self._rchol[:na * M], self._rchol[na * M:]
        Xa = ra.T.dot(self.GH[0].ravel())
        Xb = rb.T.dot(self.GH[1].ravel())
        ecoul = numpy.dot(Xa, Xa) + numpy.dot(Xb, Xb) + 2 * numpy.dot(Xa, Xb)
        Ta = numpy.einsum('ipx,jp->xij', ra.reshape(na, M, -1), self.GH[0], optimize=True)
        Tb = numpy.einsum('ipx,jp->xij', rb.reshape(nb, M, -1), self.GH[1], optimize=True)
        exx = numpy.einsum('xij,xji->', Ta, Ta) + numpy.einsum('xij,xji->', Tb, Tb)
        e2b = 0.5 * (ecoul - exx)
        self.energy = e1b + e2b + system.ecore
        self.e1b, self.e2b = e1b + system.ecore, e2b
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
        raise NotImplementedError("pauxy_b200: wavefunction files need HDF5 (SURVEY.md 8f.2); "
                                  "pass trial=MultiSlater(system, (coeffs, psi)) instead")
    na, nb = system.nup, system.ndown
    wfn = numpy.zeros((1, system.nbasis, na + nb), dtype=numpy.complex128)
    I = numpy.identity(system.nbasis, dtype=numpy.complex128)
    wfn[0, :, :na] = I[:, :na]
    wfn[0, :, na:] = I[:, :nb]
    trial = MultiSlater(system, (numpy.array([1.0 + 0j]), wfn), options=options, verbose=verbose)
    trial.half_rotate(system, scomm)
    return trial

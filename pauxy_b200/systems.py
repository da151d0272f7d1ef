"""Host-side Hamiltonian container (data provider, uploaded once).

Mirrors pauxy.systems.generic.Generic (pauxy/systems/generic.py:74-166) for the
attributes the hot path reads.  The reference's unused full SVD of the
Cholesky matrix (generic.py:157) is not reproduced.
"""
import numpy


def construct_h1e_mod(chol, h1e):
    """h1e_mod = h1e - 1/2 sum_n L_n L_n^T (pauxy/systems/generic.py:202-210)."""
    nbasis = h1e.shape[-1]
    chol3 = chol.reshape((nbasis, nbasis, -1))
    v0 = 0.5 * numpy.einsum('ikn,jkn->ij', chol3, chol3, optimize='optimal')
    return numpy.array([h1e[0] - v0, h1e[1] - v0])


class Generic(object):
    """Generic (Cholesky-factorised) many-electron Hamiltonian.

    Parameters as in the reference: nelec=(nup, ndown), h1e [2,M,M],
    chol [M*M, nchol] (row p*M+q), ecore.
    """

    def __init__(self, nelec=None, h1e=None, chol=None, ecore=None, h1e_mod=None,
                 verbose=False, exact_eri=False, stochastic_ri=False, pno=False, control_variate=False):
        self.name = "Generic"
        self.verbose = verbose
        self.nup, self.ndown = nelec
        self.nelec = nelec
        self.ne = self.nup + self.ndown
        self.ecore = ecore
        h1e = numpy.asarray(h1e)
        if h1e.ndim == 2:
            h1e = numpy.array([h1e, h1e])
        self.H1 = h1e
        self.nbasis = h1e.shape[-1]
        self.chol_vecs = numpy.ascontiguousarray(chol)
        assert self.chol_vecs.shape[0] == self.nbasis * self.nbasis
        self.cplx_chol = numpy.iscomplexobj(self.chol_vecs)
        self.sparse = False
        self.nchol = self.chol_vecs.shape[-1]
        self.nfields = self.nchol
        self.h1e_mod = h1e_mod if h1e_mod is not None else construct_h1e_mod(self.chol_vecs, self.H1)
        self.hs_pot = self.chol_vecs
        self.ktwist = numpy.array([None])
        # Alternative local-energy evaluators of the reference (systems/generic.py:77-124,
        # estimators/mixed.py:415-437).  exact_eri: the half-rotated ERI contraction
        # local_energy_generic_opt (estimators/generic.py:133-150) -- the device's ERI form of the
        # exchange, same numbers as the Cholesky form to rounding.  The sampled / truncated ones
        # (stochastic RI, PNO, control variates) are not built.
        if stochastic_ri or pno or control_variate:
            raise NotImplementedError("stochastic_ri / pno / control_variate energy evaluators "
                                      "(SURVEY.md 8f.4) are not built")
        self.control_variate = False
        self.stochastic_ri = False
        self.exact_eri = bool(exact_eri)
        self.pno = False

    def hijkl(self, i, j, k, l):
        ik = i * self.nbasis + k
        jl = j * self.nbasis + l
        return numpy.dot(self.chol_vecs[ik, :], self.chol_vecs[jl, :])

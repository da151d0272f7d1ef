"""Synthetic Hamiltonians for tests and benchmarks.

`generate_hamiltonian` restates /root/reference/pauxy/utils/testing.py:6-29
(random h1e + random 8-fold symmetric ERIs, factorised with the modified
Cholesky of /root/reference/pauxy/utils/linalg.py:112-161) and consumes the
global legacy numpy stream in the same order, so `numpy.random.seed(7)`
reproduces the reference's test inputs bit for bit.

`synthetic_cholesky_hamiltonian` builds the c2..c5 shapes of BASELINE.json
directly in factorised form: the reference generator
needs an M^4 ERI tensor, which is 12.8 GB at M=200.
"""
import numpy


def modified_cholesky(M, tol=1e-6, cmax=20):
    """Pivoted incomplete Cholesky M ~= sum_n L_n L_n^dagger.

    Follows /root/reference/pauxy/utils/linalg.py:112-161.
    Returns an array of shape [nchol, dim].
    """
    assert len(M.shape) == 2
    delta = numpy.copy(M.diagonal())
    nchol_max = int(cmax * M.shape[0] ** 0.5)
    chol_vecs = numpy.zeros((nchol_max, M.shape[0]), dtype=M.dtype)
    nu = numpy.argmax(numpy.abs(delta))
    delta_max = delta[nu]
    Mapprox = numpy.zeros(M.shape[0], dtype=M.dtype)
    chol_vecs[0] = numpy.copy(M[:, nu]) / delta_max ** 0.5
    nchol = 0
    while abs(delta_max) > tol:
        Mapprox += chol_vecs[nchol] * chol_vecs[nchol].conj()
        delta = M.diagonal() - Mapprox
        nu = numpy.argmax(numpy.abs(delta))
        delta_max = numpy.abs(delta[nu])
        nchol += 1
        Munu0 = numpy.dot(chol_vecs[:nchol, nu].conj(), chol_vecs[:nchol, :])
        chol_vecs[nchol] = (M[:, nu] - Munu0) / (delta_max) ** 0.5
    return numpy.array(chol_vecs[:nchol])


def generate_hamiltonian(nmo, nelec, cplx=False, sym=8):
    """Random Hamiltonian from the GLOBAL numpy stream (seed it first).

    Follows /root/reference/pauxy/utils/testing.py:6-29.
    Returns (h1e [M,M], chol [N,M,M], enuc, eri [M^2,M^2]).
    """
    h1e = numpy.random.random((nmo, nmo))
    if cplx:
        h1e = h1e + 1j * numpy.random.random((nmo, nmo))
    eri = numpy.random.normal(scale=0.01, size=(nmo, nmo, nmo, nmo))
    if cplx:
        eri = eri + 1j * numpy.random.normal(scale=0.01, size=(nmo, nmo, nmo, nmo))
    if sym >= 4:
        eri = eri + eri.transpose(2, 3, 0, 1)
        eri = eri + eri.transpose(3, 2, 1, 0).conj()
    if sym == 8:
        eri = eri + eri.transpose(1, 0, 2, 3)
    eri = eri.transpose((0, 1, 3, 2))
    eri = eri.reshape((nmo * nmo, nmo * nmo))
    eri = numpy.dot(eri, eri.conj().T)
    chol = modified_cholesky(eri, tol=1e-3, cmax=30)
    chol = chol.reshape((-1, nmo, nmo))
    enuc = numpy.random.rand()
    return h1e, chol, enuc, eri


def synthetic_cholesky_hamiltonian(nbasis, nchol, seed, scale=0.02, h1_scale=0.05,
                                   ramp=0.05):
    """Factorised synthetic Hamiltonian of a named shape.

    L_n = s (A_n + A_n^T)/2 with A_n ~ N(0,1)^{MxM}, s = scale/sqrt(M);
    h1e = h1_scale * sym(N(0,1)) + diag(ramp * p); ecore = 0.

    Returns (h1e [M,M] float64, hs_pot [M*M, N] float64 C-order, ecore).
    """
    rs = numpy.random.RandomState(seed)
    M = nbasis
    s = scale / numpy.sqrt(M)
    hs_pot = numpy.empty((M * M, nchol), dtype=numpy.float64)
    for n in range(nchol):
        A = rs.normal(size=(M, M))
        hs_pot[:, n] = (0.5 * s * (A + A.T)).ravel()
    B = rs.normal(size=(M, M))
    h1e = h1_scale * 0.5 * (B + B.T) + numpy.diag(ramp * numpy.arange(M))
    return h1e, hs_pot, 0.0


# BASELINE.json configs: name -> (M, (na, nb), N, W, stabilise_freq).
# c1 is generated with generate_hamiltonian (N = 77 for seed 7).
CONFIGS = {
    'c1': dict(nbasis=12, nelec=(4, 4), nchol=77, nwalkers=32, stabilise_freq=10),
    'c2': dict(nbasis=24, nelec=(5, 5), nchol=120, nwalkers=1024, stabilise_freq=10),
    'c3': dict(nbasis=60, nelec=(7, 7), nchol=300, nwalkers=4096, stabilise_freq=5),
    'c4': dict(nbasis=108, nelec=(21, 21), nchol=500, nwalkers=8192, stabilise_freq=10),
    'c5': dict(nbasis=200, nelec=(40, 40), nchol=1000, nwalkers=16384, stabilise_freq=10),
}


def make_config_hamiltonian(name):
    """(h1e, hs_pot [M*M,N], ecore, nelec) for a BASELINE config name."""
    cfg = CONFIGS[name]
    M = cfg['nbasis']
    if name == 'c1':
        state = numpy.random.get_state()
        numpy.random.seed(7)
        h1e, chol, enuc, _ = generate_hamiltonian(M, cfg['nelec'], cplx=False)
        numpy.random.set_state(state)
        hs_pot = chol.reshape((-1, M * M)).T.copy()
        return h1e, hs_pot, enuc, cfg['nelec']
    seed = {'c2': 1002, 'c3': 1003, 'c4': 1004, 'c5': 1005}[name]
    h1e, hs_pot, ecore = synthetic_cholesky_hamiltonian(M, cfg['nchol'], seed)
    return h1e, hs_pot, ecore, cfg['nelec']

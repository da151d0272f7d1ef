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


def modified_cholesky(matrix, tol=1e-6, cmax=20):
    """Pivoted, incomplete Cholesky factorisation  matrix ~= sum_n v_n v_n^dagger  of a Hermitian
    positive semi-definite matrix, stopped when the largest residual diagonal drops below `tol`.

    Same algorithm AND the same floating-point evaluation order as the reference's test helper
    (/root/reference/pauxy/utils/linalg.py:112-161: accumulated diagonal of the approximation,
    residual recomputed from it, one matrix-vector product per new vector), because the synthetic
    c1 Hamiltonian must come out bit-identical to the reference's.  Returns [nvec, dim].
    """
    if matrix.ndim != 2:
        raise ValueError("modified_cholesky needs a matrix")
    dim = matrix.shape[0]
    diagonal = matrix.diagonal()
    capacity = int(cmax * dim ** 0.5)
    vectors = numpy.zeros((capacity, dim), dtype=matrix.dtype)
    covered = numpy.zeros(dim, dtype=matrix.dtype)     # diagonal of sum_n v_n v_n^dagger so far
    pivot = int(numpy.argmax(numpy.abs(diagonal)))
    residual = diagonal[pivot]
    vectors[0] = matrix[:, pivot] / residual ** 0.5
    count = 0
    while abs(residual) > tol:
        covered += vectors[count] * vectors[count].conj()
        remaining = diagonal - covered
        pivot = int(numpy.argmax(numpy.abs(remaining)))
        residual = numpy.abs(remaining[pivot])
        count += 1
        # column `pivot` of the current approximation, subtracted from the matrix column
        overlap = numpy.dot(vectors[:count, pivot].conj(), vectors[:count, :])
        vectors[count] = (matrix[:, pivot] - overlap) / residual ** 0.5
    return vectors[:count].copy()


def _random_block(shape, cplx, gaussian_scale=None):
    """One real (or complex: real part first) block from the global legacy numpy stream."""
    def draw():
        if gaussian_scale is None:
            return numpy.random.random(shape)
        return numpy.random.normal(scale=gaussian_scale, size=shape)
    block = draw()
    if cplx:
        block = block + 1j * draw()
    return block


def generate_hamiltonian(nmo, nelec, cplx=False, sym=8):
    """Random test Hamiltonian drawn from the GLOBAL numpy stream (seed it first): uniform h1e,
    Gaussian two-electron integrals symmetrised to 4- or 8-fold symmetry, made positive
    semi-definite by squaring and factorised with modified_cholesky; then one uniform for the core
    energy.  Draw order, symmetrisation order and tolerances are those of the reference's test
    helper (/root/reference/pauxy/utils/testing.py:6-29) so that numpy.random.seed(7) gives the
    inputs of the reference's own tests bit for bit.

    Returns (h1e [M,M], chol [N,M,M], enuc, eri [M^2,M^2]).
    """
    h1e = _random_block((nmo, nmo), cplx)
    eri = _random_block((nmo,) * 4, cplx, gaussian_scale=0.01)
    if sym >= 4:
        for perm, conjugate in (((2, 3, 0, 1), False), ((3, 2, 1, 0), True)):
            partner = eri.transpose(perm)
            eri = eri + (partner.conj() if conjugate else partner)
    if sym == 8:
        eri = eri + eri.transpose(1, 0, 2, 3)
    # Hermitian super-matrix M[(i,k),(l,j)], squared to make it positive semi-definite
    super_matrix = eri.transpose((0, 1, 3, 2)).reshape((nmo * nmo, nmo * nmo))
    super_matrix = numpy.dot(super_matrix, super_matrix.conj().T)
    chol = modified_cholesky(super_matrix, tol=1e-3, cmax=30).reshape((-1, nmo, nmo))
    enuc = numpy.random.rand()
    return h1e, chol, enuc, super_matrix


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

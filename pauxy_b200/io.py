"""On-disk formats of the reference, behind an OPTIONAL h5py import.

h5py is not part of the build image, so every function here raises a clear ImportError when it is
missing and the drivers keep their .npy / .npz stand-ins (walkers.py, estimators.py).  Where h5py
exists the same dataset names and layouts as the reference are read and written:

  * QMCPACK-style Hamiltonians, dense and sparse Cholesky factors
    (pauxy/utils/io.py:129-214: 'Hamiltonian/{Energies,dims,hcore,DenseFactorized/L,
    Factorized/{block_sizes,index_%i,vals_%i}}'; complex data stored as [..., 2] doubles);
  * QMCPACK-style wavefunctions, non-orthogonal (NOMSD, CSR determinants) and particle-hole
    (PHMSD, occupation lists) expansions (utils/io.py:325-405,406-540:
    'Wavefunction/{NOMSD,PHMSD}/{dims,ci_coeffs,Psi0_alpha,Psi0_beta,PsiT_%d/...,occs}');
  * estimator output 'basic/headers', 'basic/energies/%09d', 'back_propagated/...', 'metadata'
    (estimators/utils.py:296-325, estimators/mixed.py:368-371, estimators/handler.py:118-121);
  * walker restart files, one dataset 'walker_%d' per global walker index holding
    [weight, phase, ot, phi.ravel()] (walkers/handler.py:148-161,432-485).
"""
import json

import numpy
import scipy.sparse


def _h5py():
    try:
        import h5py
    except ImportError as e:       # pragma: no cover - depends on the image
        raise ImportError("pauxy_b200.io: this file format needs h5py, which is not installed; the "
                          "drivers fall back to .npy / .npz containers (walkers.write_file, "
                          "Estimators.dump_datasets)") from e
    return h5py


def have_h5py():
    try:
        _h5py()
        return True
    except ImportError:
        return False


def to_qmcpack_complex(array):
    """complex array [...] -> float64 [..., 2] (utils/io.py:117-123)."""
    a = numpy.ascontiguousarray(array, dtype=numpy.complex128)
    return a.view(numpy.float64).reshape(a.shape + (2,))


def from_qmcpack_complex(data, shape):
    return numpy.ascontiguousarray(data, dtype=numpy.float64).view(numpy.complex128).ravel().reshape(shape)


# ------------------------------------------------------------------ Hamiltonians
def write_qmcpack_dense(hcore, chol, nelec, nmo, enuc=0.0, filename='hamiltonian.h5', real_chol=True):
    """utils/io.py:166-184.  chol: [nmo*nmo, nchol]."""
    h5py = _h5py()
    chol = numpy.asarray(chol)
    assert chol.ndim == 2 and chol.shape[0] == nmo * nmo
    with h5py.File(filename, 'w') as fh5:
        fh5['Hamiltonian/Energies'] = numpy.array([enuc, 0])
        if real_chol:
            fh5['Hamiltonian/hcore'] = numpy.real(hcore)
            fh5['Hamiltonian/DenseFactorized/L'] = numpy.real(chol)
        else:
            fh5['Hamiltonian/hcore'] = to_qmcpack_complex(hcore)
            fh5['Hamiltonian/DenseFactorized/L'] = to_qmcpack_complex(chol)
        fh5['Hamiltonian/dims'] = numpy.array([0, 0, 0, nmo, nelec[0], nelec[1], 0, chol.shape[-1]])


def read_qmcpack_dense(filename):
    """utils/io.py:186-208 -> (hcore [M,M], chol [M*M, N], enuc, nmo, nalpha, nbeta)."""
    h5py = _h5py()
    with h5py.File(filename, 'r') as fh5:
        enuc = numpy.asarray(fh5['Hamiltonian/Energies'][:])[0]
        dims = numpy.asarray(fh5['Hamiltonian/dims'][:])
        nmo, nchol = int(dims[3]), int(dims[-1])
        hcore = numpy.asarray(fh5['Hamiltonian/hcore'][:])
        chol = numpy.asarray(fh5['Hamiltonian/DenseFactorized/L'][:])
        if hcore.ndim == 3 or hcore.size == 2 * nmo * nmo:      # complex storage [..., 2]
            hcore = from_qmcpack_complex(hcore, (nmo, nmo))
            chol = from_qmcpack_complex(chol, (nmo * nmo, nchol))
        return hcore, chol.reshape(nmo * nmo, nchol), enuc, nmo, int(dims[4]), int(dims[5])


def read_qmcpack_sparse(filename):
    """utils/io.py:129-164: blocked COO Cholesky factor; returned dense [M*M, N] (the device path
    works on the dense, fragment-major form)."""
    h5py = _h5py()
    with h5py.File(filename, 'r') as fh5:
        enuc = numpy.asarray(fh5['Hamiltonian/Energies'][:])[0]
        dims = numpy.asarray(fh5['Hamiltonian/dims'][:])
        nmo, nchol = int(dims[3]), int(dims[7])
        hcore = numpy.asarray(fh5['Hamiltonian/hcore'][:])
        real_ints = hcore.size == nmo * nmo
        hcore = hcore.reshape(nmo, nmo) if real_ints else from_qmcpack_complex(hcore, (nmo, nmo))
        block_sizes = numpy.asarray(fh5['Hamiltonian/Factorized/block_sizes'][:])
        rows, cols, vals = [], [], []
        for ic, bs in enumerate(block_sizes):
            ixs = numpy.asarray(fh5['Hamiltonian/Factorized/index_%i' % ic][:])
            rows.append(ixs[::2])
            cols.append(ixs[1::2])
            v = numpy.asarray(fh5['Hamiltonian/Factorized/vals_%i' % ic][:])
            vals.append(numpy.real(v).ravel() if real_ints else from_qmcpack_complex(v, (int(bs),)))
        chol = scipy.sparse.csr_matrix((numpy.concatenate(vals), (numpy.concatenate(rows), numpy.concatenate(cols))),
                                       shape=(nmo * nmo, nchol)).toarray()
        return hcore, chol, enuc, nmo, int(dims[4]), int(dims[5])


def read_qmcpack_hamiltonian(filename):
    """Dense if the file has a DenseFactorized group, else the sparse format
    (pauxy/systems/utils.py get_generic_integrals)."""
    h5py = _h5py()
    with h5py.File(filename, 'r') as fh5:
        dense = 'Hamiltonian/DenseFactorized/L' in fh5
    return read_qmcpack_dense(filename) if dense else read_qmcpack_sparse(filename)


def system_from_file(filename, nelec=None):
    """systems.Generic from a QMCPACK Hamiltonian file (pauxy/systems/utils.py:10-60)."""
    from .systems import Generic
    hcore, chol, enuc, nmo, na, nb = read_qmcpack_hamiltonian(filename)
    if numpy.abs(numpy.imag(chol)).max() == 0.0:
        chol = numpy.real(chol)
    nelec = (na, nb) if nelec is None else tuple(nelec)
    return Generic(nelec=nelec, h1e=numpy.array([hcore, hcore]), chol=numpy.ascontiguousarray(chol), ecore=enuc)


# ------------------------------------------------------------------ wavefunctions
def write_qmcpack_wfn(filename, wfn, walker_type, nelec, norb, init=None, mode='w'):
    """utils/io.py:406-540.  wfn = (coeffs, psi[ndets, M, ne]) (NOMSD) or (coeffs, occa, occb) (PHMSD);
    walker_type 'rhf' or 'uhf'."""
    h5py = _h5py()
    na, nb = nelec
    uhf = walker_type == 'uhf'
    with h5py.File(filename, mode) as fh5:
        if len(wfn) == 3:
            coeffs, occa, occb = wfn
            base = 'Wavefunction/PHMSD/'
            if init is not None:
                fh5[base + 'Psi0_alpha'] = to_qmcpack_complex(init[0])
                fh5[base + 'Psi0_beta'] = to_qmcpack_complex(init[1])
            else:                 # utils/io.py write_phmsd: RHF reference orbitals
                eye = numpy.eye(norb, dtype=numpy.complex128)
                fh5[base + 'Psi0_alpha'] = to_qmcpack_complex(eye[:, list(occa[0])])
                fh5[base + 'Psi0_beta'] = to_qmcpack_complex(eye[:, list(occb[0])])
            occs = numpy.concatenate([numpy.asarray(occa), numpy.asarray(occb) + norb], axis=1)
            fh5[base + 'occs'] = numpy.asarray(occs, dtype=numpy.int32).ravel()
            wtype = 2
        else:
            coeffs, psi = wfn
            psi = numpy.array(psi, dtype=numpy.complex128)
            if psi.ndim == 2:
                psi = psi[None]
            base = 'Wavefunction/NOMSD/'
            first = init if init is not None else (psi[0][:, :na], psi[0][:, na:])
            fh5[base + 'Psi0_alpha'] = to_qmcpack_complex(first[0])
            if uhf:
                fh5[base + 'Psi0_beta'] = to_qmcpack_complex(first[1])
            for idet, w in enumerate(psi):
                blocks = [(2 * idet if uhf else idet, w[:, :na])]
                if uhf:
                    blocks.append((2 * idet + 1, w[:, na:]))
                for ix, orb in blocks:      # QMCPACK keeps psi^dagger as CSR
                    m = scipy.sparse.csr_matrix(orb.conj().T)
                    g = base + 'PsiT_%d/' % ix
                    fh5[g + 'dims'] = numpy.array([m.shape[0], m.shape[1], m.nnz], dtype=numpy.int32)
                    fh5[g + 'data_'] = to_qmcpack_complex(m.data)
                    fh5[g + 'jdata_'] = m.indices
                    fh5[g + 'pointers_begin_'] = m.indptr[:-1]
                    fh5[g + 'pointers_end_'] = m.indptr[1:]
            wtype = 2 if uhf else 1
        fh5[base + 'ci_coeffs'] = to_qmcpack_complex(numpy.asarray(coeffs))
        fh5[base + 'dims'] = numpy.array([norb, na, nb, wtype, len(coeffs)], dtype=numpy.int32)


def read_qmcpack_wfn(filename, nelec=None):
    """utils/io.py:325-405 -> (wfn tuple for trial.MultiSlater, psi0 [M, ne])."""
    h5py = _h5py()
    with h5py.File(filename, 'r') as fh5:
        nomsd = 'Wavefunction/NOMSD/dims' in fh5
        base = 'Wavefunction/NOMSD/' if nomsd else 'Wavefunction/PHMSD/'
        if not nomsd and 'Wavefunction/PHMSD/dims' not in fh5:
            raise KeyError("no Wavefunction/NOMSD or Wavefunction/PHMSD group in %s" % filename)
        dims = numpy.asarray(fh5[base + 'dims'][:])
        nmo, na, nb, wtype, nci = (int(x) for x in dims[:5])
        if nelec is not None and (na, nb) != tuple(nelec):
            raise ValueError("number of electrons does not match the wavefunction: %s vs %s" % ((na, nb), nelec))
        uhf = wtype == 2
        coeffs = from_qmcpack_complex(fh5[base + 'ci_coeffs'][:], (nci,))
        psi0 = numpy.zeros((nmo, na + nb), dtype=numpy.complex128)
        psi0[:, :na] = from_qmcpack_complex(fh5[base + 'Psi0_alpha'][:], (nmo, na))
        if uhf and (base + 'Psi0_beta') in fh5:
            psi0[:, na:] = from_qmcpack_complex(fh5[base + 'Psi0_beta'][:], (nmo, nb))
        else:
            psi0[:, na:] = psi0[:, :nb]
        if not nomsd:
            occs = numpy.asarray(fh5[base + 'occs'][:]).reshape((nci, na + nb))
            return (coeffs, occs[:, :na], occs[:, na:] - nmo), psi0

        def orbs(ix):
            g = base + 'PsiT_%d/' % ix
            d = numpy.asarray(fh5[g + 'dims'][:])
            data = from_qmcpack_complex(fh5[g + 'data_'][:], (int(d[2]),))
            indptr = numpy.append(numpy.asarray(fh5[g + 'pointers_begin_'][:]), int(d[2]))
            m = scipy.sparse.csr_matrix((data, numpy.asarray(fh5[g + 'jdata_'][:]), indptr),
                                        shape=(int(d[0]), int(d[1])))
            return m.toarray().conj().T
        psi = numpy.zeros((nci, nmo, na + nb), dtype=numpy.complex128)
        for idet in range(nci):
            pa = orbs(2 * idet if uhf else idet)
            psi[idet, :, :na] = pa
            psi[idet, :, na:] = orbs(2 * idet + 1) if uhf else pa[:, :nb]
        return (coeffs, psi), psi0


# ------------------------------------------------------------------ estimator output
class H5EstimatorHelper(object):
    """estimators/utils.py:296-325: one dataset per push, zero-padded running index."""

    def __init__(self, filename, base, nav=1):
        self.filename, self.base, self.index, self.nzero, self.nav = filename, base, 0, 9, nav

    def push(self, data, name):
        h5py = _h5py()
        ix = str(self.index)
        with h5py.File(self.filename, 'a') as fh5:
            fh5[self.base + '/' + name + '/' + '0' * (self.nzero - len(ix)) + ix] = data

    def increment(self):
        self.index = (self.index + 1) // self.nav

    def reset(self):
        self.index = 0


def write_estimates(filename, headers, rows, metadata=None, back_propagated=None):
    """The reference's estimates.<n>.h5 from the rows the driver collected: 'basic/headers',
    'basic/energies/%09d', 'back_propagated/{denominator,one_rdm}_<ix>/%09d', 'metadata'."""
    h5py = _h5py()
    with h5py.File(filename, 'w') as fh5:
        fh5['basic/headers'] = numpy.array(headers).astype('S')
        fh5['metadata'] = json.dumps(metadata or {})
    out = H5EstimatorHelper(filename, 'basic')
    for row in rows:
        out.push(numpy.asarray(row), 'energies')
        out.increment()
    for kind, per_ix in (back_propagated or {}).items():
        for ix, vals in per_ix.items():
            h = H5EstimatorHelper(filename, 'back_propagated')
            for v in vals:
                h.push(numpy.asarray(v), '%s_%d' % (kind, ix))
                h.increment()


# ------------------------------------------------------------------ walker restart
def write_walkers_h5(filename, buffers, first_global_index, create=True):
    """walkers/handler.py:148-161,443-454: dataset 'walker_%d' = [weight, phase, ot, phi.ravel()]
    for every global walker index owned by this rank."""
    h5py = _h5py()
    with h5py.File(filename, 'w' if create else 'a') as fh5:
        for i, buff in enumerate(buffers):
            fh5['walker_%d' % (first_global_index + i)] = numpy.asarray(buff, dtype=numpy.complex128)


def read_walkers_h5(filename, first_global_index, nwalkers):
    h5py = _h5py()
    with h5py.File(filename, 'r') as fh5:
        return numpy.array([numpy.asarray(fh5['walker_%d' % (first_global_index + i)][:])
                            for i in range(nwalkers)])

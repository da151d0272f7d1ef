"""TEST INFRASTRUCTURE -- drives the UNMODIFIED reference (/root/reference/pauxy).

Runs only in the build container (the reference tree does not exist on the GPU
box).  It is used to (1) check the reference's own goldens reproduce here,
(2) validate oracle/afqmc_oracle.py step by step and (3) generate the small
fixtures committed under tests/golden/ (see oracle/gen_golden.py).

Nothing under pauxy_b200/ imports this module.

How the reference is made importable: stub `h5py` and
`mpi4py` packages (oracle/stubs) are put on sys.path; the reference files are
not touched.  `numpy.linalg.svd` is patched out while `Generic` is constructed
because /root/reference/pauxy/systems/generic.py:157 runs an unused full SVD of
the Cholesky matrix (minutes / GBs at M >= 60).
"""
import os
import sys
from unittest import mock

import numpy

_HERE = os.path.dirname(os.path.abspath(__file__))
REFERENCE_ROOT = os.environ.get('PAUXY_REFERENCE_ROOT', '/root/reference')


def reference_available():
    return os.path.isdir(os.path.join(REFERENCE_ROOT, 'pauxy'))


def _install_paths():
    stubs = os.path.join(_HERE, 'stubs')
    for p in (REFERENCE_ROOT, stubs):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.dont_write_bytecode = True


def build_reference_afqmc(h1e, hs_pot, ecore, nelec, options, trial_factory=None):
    """Construct the reference AFQMC driver on a given Hamiltonian.

    Mirrors /root/reference/pauxy/qmc/tests/test_afqmc.py:211-217.
    """
    _install_paths()
    from mpi4py import MPI
    from pauxy.qmc.afqmc import AFQMC
    from pauxy.systems.generic import Generic
    fake_svd = lambda *a, **k: (None, None, None)
    with mock.patch('numpy.linalg.svd', fake_svd):
        system = Generic(nelec=nelec, h1e=numpy.array([h1e, h1e]),
                         chol=hs_pot, ecore=ecore)
    comm = MPI.COMM_WORLD
    if trial_factory is not None:       # e.g. a multi-determinant MultiSlater built on `system`
        afqmc = AFQMC(comm=comm, system=system, options=options, trial=trial_factory(system))
    else:
        afqmc = AFQMC(comm=comm, system=system, options=options)
    return afqmc, comm


def run_reference_untraced(h1e, hs_pot, ecore, nelec, options):
    """afqmc.run() exactly as the reference's driver test does; returns the
    driver (estimator rows are collected from the stub h5py store)."""
    afqmc, comm = build_reference_afqmc(h1e, hs_pot, ecore, nelec, options)
    afqmc.run(comm=comm, verbose=0)
    afqmc.finalise(verbose=0)
    return afqmc


def estimator_rows(filename='estimates.0.h5'):
    """Rows pushed by H5EstimatorHelper (estimators/utils.py:308-320)."""
    import h5py
    g = h5py._STORE[filename]
    keys = sorted(k for k in g.keys() if k.startswith('basic/energies/'))
    return numpy.array([numpy.asarray(g[k]) for k in keys])


def estimator_one_rdm(filename='estimates.0.h5'):
    """'basic/one_rdm/%09d' pushed by Mixed.print_step (estimators/mixed.py:279-283)."""
    import h5py
    g = h5py._STORE[filename]
    keys = sorted(k for k in g.keys() if k.startswith('basic/one_rdm/'))
    return numpy.array([numpy.asarray(g[k]) for k in keys])


def run_reference_traced(h1e, hs_pot, ecore, nelec, options, nsteps_total=None, trial_factory=None):
    """Re-run the loop body of AFQMC.run (/root/reference/pauxy/qmc/afqmc.py:
    200-255) calling the reference's own objects, recording per-step vectors.

    Returns (afqmc, trace dict of numpy arrays).
    """
    afqmc, comm = build_reference_afqmc(h1e, hs_pot, ecore, nelec, options, trial_factory)
    psi = afqmc.psi
    qmc = afqmc.qmc
    prop = afqmc.propagators
    mixed = afqmc.estimators.estimators['mixed']
    system, trial = afqmc.system, afqmc.trial
    W = len(psi.walkers)
    N = system.nfields
    total = qmc.total_steps if nsteps_total is None else nsteps_total

    tr = {k: [] for k in ('xi', 'active', 'weight_prop', 'weight', 'unscaled_weight',
                          'ot', 'hybrid_energy', 'eloc', 'parent_ix', 'comb_r',
                          'eshift', 'estimates', 'detR', 'total_weight', 'phase')}

    # record fields and the comb's random number through the real RNG calls
    real_normal = numpy.random.normal
    real_random = numpy.random.random
    step_xi = []
    step_r = []

    def rec_normal(*a, **k):
        v = real_normal(*a, **k)
        step_xi.append(numpy.array(v, copy=True))
        return v

    def rec_random(*a, **k):
        v = real_random(*a, **k)
        step_r.append(v)
        return v

    real_bcast = comm.bcast
    step_parent = []

    def rec_bcast(obj, root=0):
        if isinstance(obj, dict) and 'ix' in obj:
            step_parent.append(numpy.array(obj['ix'], copy=True))
        return real_bcast(obj, root=root)

    comm.bcast = rec_bcast
    # back-propagated estimates as BackPropagation.print_step reduces them
    # (estimators/back_propagation.py:282-333), captured before they are zeroed
    bp = afqmc.estimators.estimators.get('back_prop')
    bp_rec = []
    if bp is not None:
        real_bp_print = bp.print_step

        def rec_bp_print(comm_, nprocs, step, nsteps=1, free_projection=False):
            if bp.accumulated:
                bp_rec.append((int(bp.buff_ix), numpy.array(bp.estimates, copy=True)))
            return real_bp_print(comm_, nprocs, step, nsteps, free_projection)
        bp.print_step = rec_bp_print
    eloc_now = numpy.zeros((W, 3), dtype=numpy.complex128)
    real_local_energy = [w.local_energy for w in psi.walkers]

    try:
        numpy.random.normal = rec_normal
        numpy.random.random = rec_random
        afqmc.setup_timers()
        eshift = 0
        mixed.update(system, qmc, trial, psi, 0, prop.free_projection)
        tr['init_ot'] = numpy.array([w.ot for w in psi.walkers])
        tr['init_estimates'] = numpy.array(mixed.estimates, copy=True)
        for step in range(1, total + 1):
            del step_xi[:], step_r[:], step_parent[:]
            if step % qmc.nstblz == 0:
                psi.orthogonalise(trial, prop.free_projection)
            active = numpy.zeros(W, dtype=bool)
            for iw, w in enumerate(psi.walkers):
                if abs(w.weight) > 1e-8:
                    active[iw] = True
                    prop.propagate_walker(w, system, trial, eshift)
                if (abs(w.weight) > w.total_weight * 0.10) and step > 1:
                    w.weight = w.total_weight * 0.10
            xi = numpy.zeros((W, N))
            xi[active] = numpy.array(step_xi).reshape(-1, N)
            tr['xi'].append(xi)
            tr['active'].append(active)
            tr['weight_prop'].append(numpy.array([w.weight for w in psi.walkers]))
            if step % qmc.npop_control == 0:
                psi.pop_control(comm)
            tr['parent_ix'].append(step_parent[0] if step_parent
                                   else numpy.ones(W, dtype='i'))
            tr['comb_r'].append(step_r[0] if step_r else -1.0)
            # local energies as Mixed.update computes them (fresh Green's fn)
            if step % mixed.energy_eval_freq == 0:
                for iw, w in enumerate(psi.walkers):
                    w.greens_function(trial)
                    eloc_now[iw] = w.local_energy(system, rchol=trial._rchol,
                                                  eri=trial._eri, UVT=trial._UVT)
            afqmc.estimators.update(system, qmc, trial, psi, step,
                                    prop.free_projection)
            tr['estimates'].append(numpy.array(mixed.estimates, copy=True))
            afqmc.estimators.print_step(comm, comm.size, step)
            if step < qmc.neqlb:
                eshift = mixed.get_shift(prop.hybrid)
            else:
                eshift += (mixed.get_shift() - eshift)
            tr['eshift'].append(eshift)
            tr['weight'].append(numpy.array([w.weight for w in psi.walkers]))
            tr['unscaled_weight'].append(numpy.array([w.unscaled_weight for w in psi.walkers]))
            tr['ot'].append(numpy.array([w.ot for w in psi.walkers], dtype=numpy.complex128))
            tr['hybrid_energy'].append(numpy.array([w.hybrid_energy for w in psi.walkers],
                                                   dtype=numpy.complex128))
            tr['detR'].append(numpy.array([w.detR for w in psi.walkers]))
            tr['phase'].append(numpy.array([w.phase for w in psi.walkers], dtype=numpy.complex128))
            tr['total_weight'].append(psi.walkers[0].total_weight)
            tr['eloc'].append(eloc_now.copy())
    finally:
        numpy.random.normal = real_normal
        numpy.random.random = real_random
        comm.bcast = real_bcast
    out = {k: numpy.array(v) for k, v in tr.items()}
    out['phi_final'] = numpy.array([w.phi for w in psi.walkers])
    out['nfb_trig'] = numpy.array(prop.nfb_trig)
    out['nhe_trig'] = numpy.array(prop.nhe_trig)
    out['rows'] = estimator_rows(afqmc.estimators.filename)
    if mixed.calc_one_rdm:
        out['mixed_one_rdm'] = estimator_one_rdm(afqmc.estimators.filename)
    if bp is not None:
        M = system.nbasis
        out['bp_buff_ix'] = numpy.array([b[0] for b in bp_rec])
        out['bp_energies'] = numpy.array([b[1][:bp.nreg] for b in bp_rec])
        out['bp_denominator'] = numpy.array([b[1][bp.nreg] for b in bp_rec])
        out['bp_one_rdm'] = numpy.array([b[1][bp.nreg + 1:bp.nreg + 1 + 2 * M * M].reshape(2, M, M)
                                         for b in bp_rec])
        out['phi_old_final'] = numpy.array([w.phi_old for w in psi.walkers])
    return afqmc, out


def reference_setup_arrays(afqmc):
    """Arrays the reference derives at construction (for checking the host
    setup of the product and the oracle restatement)."""
    p = afqmc.propagators.propagator
    t = afqmc.trial
    return dict(mf_shift=numpy.array(p.mf_shift), BH1=numpy.array(p.BH1),
                mf_core=numpy.array(p.mf_core), rchol=numpy.array(t._rchol),
                psi=numpy.array(t.psi), h1e_mod=numpy.array(afqmc.system.h1e_mod))

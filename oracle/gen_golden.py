"""TEST INFRASTRUCTURE -- regenerates tests/golden/*.npz from the UNMODIFIED
reference (run in the build container only):

    python oracle/gen_golden.py

Each fixture holds the inputs needed to rebuild the case (or the seeds that
regenerate them) and per-step vectors recorded from the reference's own
objects by oracle/ref_harness.py.
"""
import copy
import os
import sys

import numpy

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_harness as rh  # noqa: E402
from pauxy_b200.hamiltonians import (generate_hamiltonian,  # noqa: E402
                                     synthetic_cholesky_hamiltonian)

GOLD = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                    'tests', 'golden')


def options(nwalkers, dt, steps, blocks, seed, stab=10, popc=1, walkers=None):
    o = {'verbosity': 0, 'get_sha1': False,
         'qmc': {'timestep': dt, 'steps': steps, 'blocks': blocks, 'rng_seed': seed,
                 'num_walkers': nwalkers, 'stabilise_freq': stab,
                 'pop_control_freq': popc},
         'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}},
         'trial': {'name': 'MultiSlater'}}
    if walkers:
        o['walkers'] = walkers
    return o


def save(name, meta, tr, setup=None, keep_xi=True, keep_phi=True, extra=None):
    out = dict(meta)
    if 'phase' in tr:
        out['phase'] = tr['phase']
    for k in ('weight_prop', 'weight', 'unscaled_weight', 'ot', 'hybrid_energy', 'eloc',
              'parent_ix', 'comb_r', 'eshift', 'estimates', 'detR', 'total_weight',
              'init_ot', 'init_estimates', 'nfb_trig', 'nhe_trig', 'rows', 'active'):
        out[k] = tr[k]
    if keep_xi:
        out['xi'] = tr['xi']
    if keep_phi:
        out['phi_final'] = tr['phi_final']
    else:
        out['phi_final_head'] = tr['phi_final'][:2]
    if setup is not None:
        for k, v in setup.items():
            out['setup_' + k] = v
    if extra:
        out.update(extra)
    path = os.path.join(GOLD, name + '.npz')
    numpy.savez_compressed(path, **out)
    print(name, '%.1f KB' % (os.path.getsize(path) / 1024.),
          'nfb', int(tr['nfb_trig']), 'nhe', int(tr['nhe_trig']),
          'comb events', int((tr['parent_ix'] != 1).sum()),
          'min/max w_prop %.3g %.3g' % (tr['weight_prop'].min(), tr['weight_prop'].max()))


def case_test_generic():
    """pauxy/qmc/tests/test_afqmc.py:190-229 verbatim inputs."""
    numpy.random.seed(7)
    h1e, chol, enuc, _ = generate_hamiltonian(11, (3, 3), cplx=False)
    hs = chol.reshape((-1, 121)).T.copy()
    opts = {'verbosity': 0, 'get_sha1': False,
            'qmc': {'timestep': 0.005, 'steps': 10, 'blocks': 10, 'rng_seed': 8},
            'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}},
            'trial': {'name': 'MultiSlater'}}
    a = rh.run_reference_untraced(h1e, hs, enuc, (3, 3), copy.deepcopy(opts))
    m = a.estimators.estimators['mixed']
    m.update(a.system, a.qmc, a.trial, a.psi, 0)
    numer = m.estimates[m.names.enumer]
    rows = rh.estimator_rows()
    a2, tr = rh.run_reference_traced(h1e, hs, enuc, (3, 3), copy.deepcopy(opts))
    assert numpy.array_equal(rows[:, :10], tr['rows'][:, :10])
    meta = dict(h1e=h1e, hs_pot=hs, ecore=enuc, nelec=numpy.array((3, 3)), dt=0.005,
                nwalkers=10, steps=10, blocks=10, seed=8, stab=10, popc=1,
                ref_numer_after_run=numer, ref_test_golden_numer=3.8763193646854273,
                ref_test_golden_etotal=1.5485077038208)
    save('test_generic', meta, tr, setup=rh.reference_setup_arrays(a2))


def case_c1():
    numpy.random.seed(7)
    h1e, chol, enuc, _ = generate_hamiltonian(12, (4, 4), cplx=False)
    hs = chol.reshape((-1, 144)).T.copy()
    opts = options(32, 0.005, 10, 4, 8, stab=5, popc=1)
    a, tr = rh.run_reference_traced(h1e, hs, enuc, (4, 4), opts)
    meta = dict(h1e=h1e, hs_pot=hs, ecore=enuc, nelec=numpy.array((4, 4)), dt=0.005,
                nwalkers=32, steps=10, blocks=4, seed=8, stab=5, popc=1)
    save('c1', meta, tr, setup=rh.reference_setup_arrays(a))


def case_free(name, free, force_bias, pop='comb', walkers=None):
    """Free projection (propagation/continuous.py:175-200, walkers/handler.py:178-181,
    estimators/mixed.py:151-177; the reference switches the force bias off there,
    continuous.py:30-33) and the phaseless walk without force bias (continuous.py:136-138)."""
    numpy.random.seed(7)
    h1e, chol, enuc, _ = generate_hamiltonian(12, (4, 4), cplx=False)
    hs = 4.0 * chol.reshape((-1, 144)).T.copy()   # scaled: comb / pair-branch events must fire
    opts = options(16, 0.01, 5, 6, 8, stab=4, popc=1, walkers=walkers)
    opts['propagator'] = {'free_projection': free, 'force_bias': force_bias}
    a, tr = rh.run_reference_traced(h1e, hs, enuc, (4, 4), opts)
    meta = dict(h1e=h1e, hs_pot=hs, ecore=enuc, nelec=numpy.array((4, 4)), dt=0.01,
                nwalkers=16, steps=5, blocks=6, seed=8, stab=4, popc=1,
                free_projection=free, force_bias=force_bias, pop_control=pop,
                min_weight=(walkers or {}).get('min_weight', 0.1),
                max_weight=(walkers or {}).get('max_weight', 4.0))
    save(name, meta, tr, setup=rh.reference_setup_arrays(a))


def case_bp(name, nmo, nelec, nwalkers, tau_bp, nsplit, stab, scale, dt=0.005, steps=10, blocks=3,
            restore_weights=None):
    """Back propagation (estimators/back_propagation.py:127-225, propagation/generic.py:253-290,
    walkers/stack.py:5-127).  'bp_ref' is the reference's own driver test
    (qmc/tests/test_afqmc.py:232-278); 'bp_stress' re-orthogonalises inside the back
    propagation, splits it in two and has comb events that move field histories."""
    numpy.random.seed(7)
    h1e, chol, enuc, _ = generate_hamiltonian(nmo, nelec, cplx=False)
    hs = scale * chol.reshape((-1, nmo * nmo)).T.copy()
    opts = options(nwalkers, dt, steps, blocks, 8, stab=stab, popc=1)
    bpo = {'tau_bp': tau_bp, 'one_rdm': True, 'nsplit': nsplit}
    if restore_weights is not None:
        bpo['restore_weights'] = restore_weights     # back_propagation.py:75-80,187-196
    opts['estimator'] = {'back_propagated': bpo,
                         'mixed': {'energy_eval_freq': 1, 'verbose': False}}
    opts.pop('estimates', None)
    a, tr = rh.run_reference_traced(h1e, hs, enuc, nelec, opts)
    meta = dict(h1e=h1e, hs_pot=hs, ecore=enuc, nelec=numpy.array(nelec), dt=dt,
                nwalkers=nwalkers, steps=steps, blocks=blocks, seed=8, stab=stab, popc=1,
                tau_bp=tau_bp, nsplit=nsplit, restore_weights=str(restore_weights))
    extra = {k: tr[k] for k in ('bp_buff_ix', 'bp_denominator', 'bp_one_rdm', 'phi_old_final')}
    print(name, 'bp prints', len(tr['bp_buff_ix']), 'buff_ix', tr['bp_buff_ix'][:6])
    save(name, meta, tr, setup=rh.reference_setup_arrays(a), extra=extra)


def case_mixed_rdm():
    """Mixed one-body density matrix (estimators/mixed.py:226-229,279-283) on the stress walk."""
    numpy.random.seed(11)
    h1e, chol, enuc, _ = generate_hamiltonian(8, (3, 3), cplx=False)
    hs = 6.0 * chol.reshape((-1, 64)).T.copy()
    opts = options(16, 0.02, 5, 4, 21, stab=3, popc=1)
    opts['estimates']['mixed']['one_rdm'] = True
    a, tr = rh.run_reference_traced(h1e, hs, enuc, (3, 3), opts)
    meta = dict(h1e=h1e, hs_pot=hs, ecore=enuc, nelec=numpy.array((3, 3)), dt=0.02,
                nwalkers=16, steps=5, blocks=4, seed=21, stab=3, popc=1)
    save('mixed_rdm', meta, tr, setup=rh.reference_setup_arrays(a),
         extra={'mixed_one_rdm': tr['mixed_one_rdm']})


def case_stress(name, pop, walkers=None, scale_chol=6.0, dt=0.02, nwalkers=16, keep_xi=True):
    """Small case scaled so that the force-bias clip, the hybrid-energy bound,
    the weight cap and comb/pair-branch events all fire.  nwalkers=64 gives the multi-device
    tests and the bench's N-rank self-check clones that cross several devices."""
    numpy.random.seed(11)
    h1e, chol, enuc, _ = generate_hamiltonian(8, (3, 3), cplx=False)
    hs = scale_chol * chol.reshape((-1, 64)).T.copy()
    opts = options(nwalkers, dt, 5, 6, 21, stab=3, popc=1, walkers=walkers)
    a, tr = rh.run_reference_traced(h1e, hs, enuc, (3, 3), opts)
    meta = dict(h1e=h1e, hs_pot=hs, ecore=enuc, nelec=numpy.array((3, 3)), dt=dt,
                nwalkers=nwalkers, steps=5, blocks=6, seed=21, stab=3, popc=1,
                pop_control=pop, min_weight=(walkers or {}).get('min_weight', 0.1),
                max_weight=(walkers or {}).get('max_weight', 4.0))
    save(name, meta, tr, setup=rh.reference_setup_arrays(a), keep_xi=keep_xi)


def case_complex(name, nmo, nelec, nwalkers, scale_chol=1.0, dt=0.005, steps=5, blocks=3, stab=3):
    """Complex integrals (the reference's generate_hamiltonian(cplx=True, sym=4),
    systems/tests/test_generic.py:30): complex h1e (made Hermitian), complex Cholesky vectors, hence
    complex half-rotated vectors, mean-field shift and one-body propagator.  The unmodified
    reference driver runs it through the same code path as a real Hamiltonian."""
    numpy.random.seed(7)
    h1e, chol, enuc, _ = generate_hamiltonian(nmo, nelec, cplx=True, sym=4)
    h1e = 0.5 * (h1e + h1e.conj().T)
    hs = scale_chol * chol.reshape((-1, nmo * nmo)).T.copy()
    opts = options(nwalkers, dt, steps, blocks, 8, stab=stab, popc=1)
    a, tr = rh.run_reference_traced(h1e, hs, enuc, nelec, opts)
    meta = dict(h1e=h1e, hs_pot=hs, ecore=enuc, nelec=numpy.array(nelec), dt=dt, nwalkers=nwalkers,
                steps=steps, blocks=blocks, seed=8, stab=stab, popc=1)
    save(name, meta, tr, setup=rh.reference_setup_arrays(a))


def case_shape(name, M, na, N, W, steps, seed_h, stab, blocks=1, dt=0.005, scale=0.02, ramp=0.05):
    """BASELINE c2..c5 shapes at reduced walker count; inputs regenerate from seeds, fields from
    the global legacy stream.  The `*_shape` fixtures use the benchmark's benign Hamiltonian
    (no branch fires); the `*_stress` ones scale the Cholesky vectors, the orbital-energy ramp
    and the time step so that the force-bias clip, the hybrid-energy bound (armed from block 2
    on, when eshift != 0), comb events and re-orthogonalisations all occur AT those shapes."""
    h1e, hs, ecore = synthetic_cholesky_hamiltonian(M, N, seed_h, scale=scale, ramp=ramp)
    opts = options(W, dt, steps, blocks, 8, stab=stab, popc=1)
    a, tr = rh.run_reference_traced(h1e, hs, ecore, (na, na), opts)
    meta = dict(nbasis=M, nelec=numpy.array((na, na)), nchol=N, dt=dt, nwalkers=W,
                steps=steps, blocks=blocks, seed=8, stab=stab, popc=1, seed_h=seed_h,
                scale=scale, ramp=ramp, h1e_checksum=h1e.sum(), hs_checksum=hs.sum())
    save(name, meta, tr, keep_xi=False, keep_phi=False)


def _phmsd_inputs(nmo, nelec, ndet, seed=7, scale=1.0):
    """Inputs of pauxy/propagation/tests/test_generic.py:52-92: random Hamiltonian, particle-hole
    multi-determinant trial with random complex coefficients and a random complex initial walker."""
    rh._install_paths()
    from pauxy.utils.testing import get_random_phmsd
    numpy.random.seed(seed)
    h1e, chol, enuc, _ = generate_hamiltonian(nmo, nelec, cplx=False)
    hs = scale * chol.reshape((-1, nmo * nmo)).T.copy()

    class _S(object):
        nbasis, nup, ndown = nmo, nelec[0], nelec[1]
    wfn, init = get_random_phmsd(_S, ndet=ndet, init=True)
    return h1e, hs, enuc, wfn, init


def case_multi_det_walker(name, hybrid):
    """The reference's own multi-determinant propagation tests (propagation/tests/test_generic.py:
    52-70 hybrid=False -> local-energy weight update, :72-92 hybrid=True): one MultiDetWalker, ten
    propagate_walker calls with eshift = trial.energy (complex), known final weights
    0.68797524675701 / 0.7430443466368197."""
    from unittest import mock
    h1e, hs, enuc, wfn, init = _phmsd_inputs(10, (5, 5), 3)
    from pauxy.systems.generic import Generic
    from pauxy.trial_wavefunction.multi_slater import MultiSlater
    from pauxy.propagation.continuous import Continuous
    from pauxy.utils.misc import dotdict
    from pauxy.walkers.multi_det import MultiDetWalker
    system = Generic(nelec=(5, 5), h1e=numpy.array([h1e, h1e]), chol=hs, ecore=0)
    trial = MultiSlater(system, wfn, init=init)
    trial.calculate_energy(system)
    prop = Continuous(system, trial, dotdict({'dt': 0.005, 'nstblz': 5}), options={'hybrid': hybrid})
    walker = MultiDetWalker(system, trial)
    tr = {k: [] for k in ('xi', 'weight', 'ot', 'hybrid_energy', 'eloc', 'ovlps')}
    tr_init_ot = walker.ot
    real_normal = numpy.random.normal

    def rec(*a, **k):
        v = real_normal(*a, **k)
        tr['xi'].append(numpy.array(v, copy=True))
        return v
    numpy.random.normal = rec
    try:
        for i in range(10):
            prop.propagate_walker(walker, system, trial, trial.energy)
            tr['weight'].append(walker.weight)
            tr['ot'].append(walker.ot)
            tr['hybrid_energy'].append(walker.hybrid_energy)
            tr['eloc'].append(walker.eloc)
            tr['ovlps'].append(walker.ovlps.copy())
    finally:
        numpy.random.normal = real_normal
    out = dict(h1e=h1e, hs_pot=hs, ecore=0.0, nelec=numpy.array((5, 5)), dt=0.005, hybrid=hybrid,
               coeffs=numpy.array(wfn[0]), occa=numpy.array(wfn[1]), occb=numpy.array(wfn[2]),
               init=init, trial_energy=numpy.array([trial.energy, trial.e1b, trial.e2b]),
               init_ot=tr_init_ot, mf_shift=numpy.array(prop.propagator.mf_shift),
               BH1=numpy.array(prop.propagator.BH1),
               ref_test_golden_weight=0.7430443466368197 if hybrid else 0.68797524675701)
    out.update({k: numpy.array(v) for k, v in tr.items()})
    path = os.path.join(GOLD, name + '.npz')
    numpy.savez_compressed(path, **out)
    print(name, 'final weight %.16g' % walker.weight, 'golden', out['ref_test_golden_weight'])


def case_multi_det_driver(name, hybrid=True):
    """Full driver run (comb, re-orthogonalisation, block output) with a 3-determinant
    particle-hole trial: walkers/handler.py:64-70 picks MultiDetWalker, estimators/mixed.py:211-221
    evaluates local_energy_multi_det."""
    h1e, hs, enuc, wfn, init = _phmsd_inputs(10, (5, 5), 3, scale=3.0)

    def factory(system):
        from pauxy.trial_wavefunction.multi_slater import MultiSlater
        return MultiSlater(system, wfn, init=init)
    opts = options(16, 0.01, 5, 4, 8, stab=4, popc=1)
    opts['propagator'] = {'hybrid': hybrid}
    a, tr = rh.run_reference_traced(h1e, hs, enuc, (5, 5), opts, trial_factory=factory)
    meta = dict(h1e=h1e, hs_pot=hs, ecore=enuc, nelec=numpy.array((5, 5)), dt=0.01,
                nwalkers=16, steps=5, blocks=4, seed=8, stab=4, popc=1, hybrid=hybrid,
                coeffs=numpy.array(wfn[0]), occa=numpy.array(wfn[1]), occb=numpy.array(wfn[2]),
                init=init)
    p = a.propagators.propagator
    save(name, meta, tr, setup=dict(mf_shift=numpy.array(p.mf_shift), BH1=numpy.array(p.BH1),
                                    mf_core=numpy.array(p.mf_core)))


def case_multi_det_nomsd_complex(name):
    """Driver run with a NON-orthogonal multi-determinant trial whose orbitals and coefficients are
    complex (utils/testing.py:31-45 get_random_nomsd(cplx=True)): complex per-determinant orbitals,
    half-rotated Cholesky vectors and one-body integrals on a real Hamiltonian."""
    rh._install_paths()
    from pauxy.utils.testing import get_random_nomsd
    numpy.random.seed(7)
    nmo, nelec = 10, (4, 3)
    h1e, chol, enuc, _ = generate_hamiltonian(nmo, nelec, cplx=False)
    hs = 2.0 * chol.reshape((-1, nmo * nmo)).T.copy()

    class _S(object):
        nbasis, nup, ndown = nmo, nelec[0], nelec[1]
    coeffs, wfn = get_random_nomsd(_S, ndet=3, cplx=True)
    # orthonormal columns per determinant and spin keep the overlaps O(1)
    for i in range(wfn.shape[0]):
        wfn[i, :, :nelec[0]] = numpy.linalg.qr(wfn[i, :, :nelec[0]])[0]
        wfn[i, :, nelec[0]:] = numpy.linalg.qr(wfn[i, :, nelec[0]:])[0]
    init = wfn[0].copy()

    def factory(system):
        from pauxy.trial_wavefunction.multi_slater import MultiSlater
        return MultiSlater(system, (coeffs, wfn), init=init)
    opts = options(12, 0.01, 5, 3, 8, stab=3, popc=1)
    a, tr = rh.run_reference_traced(h1e, hs, enuc, nelec, opts, trial_factory=factory)
    meta = dict(h1e=h1e, hs_pot=hs, ecore=enuc, nelec=numpy.array(nelec), dt=0.01,
                nwalkers=12, steps=5, blocks=3, seed=8, stab=3, popc=1, hybrid=True,
                coeffs=numpy.array(coeffs), orbitals=numpy.array(wfn), init=init)
    p = a.propagators.propagator
    save(name, meta, tr, setup=dict(mf_shift=numpy.array(p.mf_shift), BH1=numpy.array(p.BH1),
                                    mf_core=numpy.array(p.mf_core)))


def case_local_energy():
    """pauxy/estimators/tests/test_generic.py:33-64 inputs and golden."""
    rh._install_paths()
    from pauxy.systems.generic import Generic
    from pauxy.trial_wavefunction.multi_slater import MultiSlater
    from pauxy.estimators.greens_function import gab_spin
    from pauxy.estimators.generic import local_energy_generic_cholesky_opt
    from pauxy.utils.testing import get_random_nomsd
    numpy.random.seed(7)
    nmo = 24
    nelec = (4, 2)
    h1e, chol, enuc, eri = generate_hamiltonian(nmo, nelec, cplx=False)
    system = Generic(nelec=nelec, h1e=numpy.array([h1e, h1e]),
                     chol=chol.reshape((-1, nmo * nmo)).T.copy(), ecore=enuc)
    wfn = get_random_nomsd(system, ndet=1, cplx=False)
    trial = MultiSlater(system, wfn)
    trial.half_rotate(system)
    e = local_energy_generic_cholesky_opt(system, trial.G, Ghalf=trial.GH,
                                          rchol=trial._rchol)
    path = os.path.join(GOLD, 'local_energy.npz')
    numpy.savez_compressed(path, h1e=h1e, hs_pot=system.hs_pot, ecore=enuc,
                           nelec=numpy.array(nelec), psi=trial.psi[0],
                           energy=numpy.array(e),
                           ref_test_golden=numpy.array([20.6826247016273,
                                                        23.0173528796140,
                                                        -2.3347281779866]))
    print('local_energy', e)


if __name__ == '__main__':
    os.makedirs(GOLD, exist_ok=True)
    which = sys.argv[1:] or ['tg', 'c1', 'stress', 'pb', 'le', 'c2s', 'c3s', 'c4s', 'free', 'bp', 'rdm']
    if 'md' in which:
        case_multi_det_walker('md_hybrid', True)
        case_multi_det_walker('md_local_energy', False)
        case_multi_det_driver('md_driver', True)
        case_multi_det_driver('md_driver_le', False)
    if 'free' in which:
        case_free('free_comb', True, True)
        case_free('free_pair_branch', True, False, pop='pair_branch',
                  walkers={'population_control': 'pair_branch', 'min_weight': 0.9,
                           'max_weight': 1.1})
        case_free('phaseless_nofb', False, False)
    if 'rdm' in which:
        case_mixed_rdm()
    if 'bp' in which:
        case_bp('bp_ref', 11, (3, 3), 10, 0.025, 1, 10, 1.0, blocks=10)
        case_bp('bp_stress', 12, (4, 4), 16, 0.12, 2, 2, 6.0, dt=0.02, steps=5, blocks=4)
    if 'bpw' in which:
        case_bp('bp_restore_full', 12, (4, 4), 16, 0.06, 2, 2, 3.0, dt=0.01, steps=5, blocks=4,
                restore_weights='full')
        case_bp('bp_restore_partial', 12, (4, 4), 16, 0.06, 2, 2, 3.0, dt=0.01, steps=5, blocks=4,
                restore_weights='partial')
    if 'tg' in which:
        case_test_generic()
    if 'c1' in which:
        case_c1()
    if 'stress' in which:
        case_stress('stress_comb', 'comb')
    if 'stress64' in which:
        case_stress('stress_comb64', 'comb', nwalkers=64, keep_xi=False)
        case_stress('stress_pair_branch64', 'pair_branch',
                    walkers={'population_control': 'pair_branch',
                             'min_weight': 0.9, 'max_weight': 1.1},
                    scale_chol=3.0, dt=0.01, nwalkers=64, keep_xi=False)
    if 'logshift' in which:
        case_stress('stress_logshift', 'comb', walkers={'use_log_shift': True})
    if 'pb' in which:
        case_stress('stress_pair_branch', 'pair_branch',
                    walkers={'population_control': 'pair_branch',
                             'min_weight': 0.9, 'max_weight': 1.1},
                    scale_chol=3.0, dt=0.01)
    if 'mdc' in which:
        case_multi_det_nomsd_complex('md_nomsd_cplx')
    if 'cplx' in which:
        case_complex('cplx_driver', 10, (3, 3), 12)
        case_complex('cplx_stress', 10, (4, 2), 16, scale_chol=4.0, dt=0.02, steps=5, blocks=4, stab=3)
    if 'le' in which:
        case_local_energy()
    if 'c2s' in which:
        case_shape('c2_shape', 24, 5, 120, 32, 12, 1002, 5)
    if 'c3s' in which:
        case_shape('c3_shape', 60, 7, 300, 16, 6, 1003, 5)
    if 'c4s' in which:
        case_shape('c4_shape', 108, 21, 500, 16, 6, 1004, 5)
    if 'c2st' in which:
        case_shape('c2_stress', 24, 5, 120, 32, 10, 1002, 5, blocks=3, dt=0.02, scale=0.4, ramp=0.2)
    if 'c3st' in which:
        case_shape('c3_stress', 60, 7, 300, 16, 10, 1003, 5, blocks=2, dt=0.02, scale=0.32, ramp=0.2)
    if 'c4st' in which:
        case_shape('c4_stress', 108, 21, 500, 32, 10, 1004, 5, blocks=2, dt=0.02, scale=0.25, ramp=0.2)
    if 'c5st' in which:
        case_shape('c5_stress', 200, 40, 1000, 8, 4, 1005, 3, blocks=2, dt=0.02, scale=0.32, ramp=0.2)
    if 'c5s' in which:
        case_shape('c5_shape', 200, 40, 1000, 4, 3, 1005, 2)

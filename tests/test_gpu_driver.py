"""Driver-level parity on the GPU: the product's AFQMC loop against vectors
recorded from the unmodified reference (tests/golden, oracle/gen_golden.py).

Bar (BASELINE.json north_star): per-step walker overlaps, weights and local
energies <= 1e-10 relative; population-control selection bit-exact."""
import numpy
import pytest

from pauxy_b200.systems import Generic
from pauxy_b200.qmc import AFQMC
from pauxy_b200.hamiltonians import synthetic_cholesky_hamiltonian

pytestmark = pytest.mark.gpu

RTOL = 1e-10
# hybrid energy E_h = -(log(ot_new / ot_old) + cfb + cmf) / dt: an overlap ratio matched to 1e-10
# relative puts <= 1e-10 / dt of absolute error on E_h; the bar used here is ten times tighter
EH_ATOL = 1e-11


def _options(g, walkers=None, propagator=None, back_propagated=None):
    o = {'qmc': {'timestep': float(g['dt']), 'steps': int(g['steps']), 'blocks': int(g['blocks']),
                 'rng_seed': int(g['seed']), 'num_walkers': int(g['nwalkers']),
                 'stabilise_freq': int(g['stab']), 'pop_control_freq': int(g['popc'])},
         'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}},
         'trial': {'name': 'MultiSlater'}}
    if walkers:
        o['walkers'] = walkers
    if propagator:
        o['propagator'] = propagator
    if back_propagated:
        o['estimates']['back_propagated'] = back_propagated
    return o


def _run(g, h1e, hs, ecore, walkers=None, propagator=None, back_propagated=None, top=None,
         trial_factory=None):
    nelec = tuple(int(x) for x in g['nelec'])
    system = Generic(nelec=nelec, h1e=numpy.array([h1e, h1e]), chol=hs, ecore=ecore)
    opts = _options(g, walkers, propagator, back_propagated)
    opts.update(top or {})
    trial = trial_factory(system) if trial_factory is not None else None
    afqmc = AFQMC(options=opts, system=system, trial=trial, verbose=0)
    hist = {k: [] for k in ('weight', 'unscaled_weight', 'ot', 'hybrid_energy', 'eloc',
                            'parent_ix', 'phase')}

    def obs(step, a):
        e = a.engine
        hist['weight'].append(e.weight.cpu().numpy().copy())
        hist['unscaled_weight'].append(e.unscaled_weight.cpu().numpy().copy())
        hist['ot'].append(e.ot.cpu().numpy().copy())
        hist['hybrid_energy'].append(e.hybrid_energy.cpu().numpy().copy())
        hist['eloc'].append(e.eloc.cpu().numpy().copy())
        hist['parent_ix'].append(e.parent_ix.cpu().numpy()[:e.W].copy())
        hist['phase'].append(e.phase.cpu().numpy().copy())
    afqmc.run(verbose=0, observer=obs)
    return afqmc, {k: numpy.array(v) for k, v in hist.items()}


def _close(a, b, rtol=RTOL, atol=0.0):
    numpy.testing.assert_allclose(a, b, rtol=rtol, atol=atol)


@pytest.mark.parametrize('name', ['test_generic', 'c1', 'stress_comb', 'cplx_driver', 'cplx_stress'])
def test_trace_matches_reference(golden, name):
    g = golden(name)
    afqmc, h = _run(g, g['h1e'], g['hs_pot'], float(g['ecore']))
    assert numpy.array_equal(h['parent_ix'], g['parent_ix'])      # bit-exact selection
    _close(h['weight'], g['weight'], atol=1e-13)
    _close(h['unscaled_weight'], g['unscaled_weight'], atol=1e-13)
    _close(h['ot'], g['ot'])
    _close(h['hybrid_energy'], g['hybrid_energy'], rtol=RTOL, atol=EH_ATOL / float(g['dt']))
    _close(h['eloc'], g['eloc'], atol=1e-10)
    assert afqmc.propagators.nfb_trig == int(g['nfb_trig'])
    assert afqmc.propagators.nhe_trig == int(g['nhe_trig'])
    rows = afqmc.estimators.rows()
    _close(rows[:, :10], g['rows'][:, :10], atol=1e-10)
    _close(afqmc.psi.phi_host(), g['phi_final'], atol=1e-11)


def test_exact_eri_option_matches_reference(golden):
    """system.exact_eri (systems/generic.py:77, estimators/mixed.py:427-428): the reference then
    contracts the half-rotated ERI built from the same Cholesky vectors (local_energy_generic_opt) --
    the same numbers as the default evaluator to rounding; here it selects the ERI form of the
    exchange kernel."""
    g = golden('c1')
    nelec = tuple(int(x) for x in g['nelec'])
    system = Generic(nelec=nelec, h1e=numpy.array([g['h1e'], g['h1e']]), chol=g['hs_pot'],
                     ecore=float(g['ecore']), exact_eri=True)
    afqmc = AFQMC(options=_options(g, None, None, None), system=system, verbose=0)
    assert afqmc.engine.exchange_is_eri()
    afqmc.run(verbose=0)
    _close(afqmc.estimators.rows()[:, :10], g['rows'][:, :10], atol=1e-10)
    with pytest.raises(NotImplementedError):
        Generic(nelec=nelec, h1e=g['h1e'], chol=g['hs_pot'], ecore=0.0, stochastic_ri=True)


@pytest.mark.parametrize('top,replays', [({}, True), ({'cuda_graphs': False}, False),
                                         ({'fused_step': False}, False)])
def test_fused_step_and_graph_replay(golden, top, replays):
    """The default driver step is ONE library call (pxb_step) replayed from a CUDA graph; with the
    graphs off, or with the phase-by-phase calls of the reference's loop body, the trace is the same."""
    g = golden('stress_comb')
    afqmc, h = _run(g, g['h1e'], g['hs_pot'], float(g['ecore']), top=top)
    assert (afqmc.engine.step_graphs() > 0) == replays
    assert numpy.array_equal(h['parent_ix'], g['parent_ix'])
    _close(h['weight'], g['weight'], atol=1e-13)
    _close(h['ot'], g['ot'])
    _close(h['hybrid_energy'], g['hybrid_energy'], rtol=RTOL, atol=EH_ATOL / float(g['dt']))
    _close(h['eloc'], g['eloc'], atol=1e-10)
    assert afqmc.propagators.nfb_trig == int(g['nfb_trig'])
    assert afqmc.propagators.nhe_trig == int(g['nhe_trig'])
    _close(afqmc.estimators.rows()[:, :10], g['rows'][:, :10], atol=1e-10)


def test_log_shift(golden):
    """walkers.use_log_shift (walkers/handler.py:228,456-475) against a trace of the reference."""
    g = golden('stress_logshift')
    afqmc, h = _run(g, g['h1e'], g['hs_pot'], float(g['ecore']), walkers={'use_log_shift': True})
    assert numpy.array_equal(h['parent_ix'], g['parent_ix'])
    _close(h['weight'], g['weight'], atol=1e-13)
    _close(h['ot'], g['ot'])
    _close(h['eloc'], g['eloc'], atol=1e-10)
    _close(afqmc.engine.detR.cpu().numpy(), g['detR'][-1], rtol=1e-10)
    _close(afqmc.estimators.rows()[:, :10], g['rows'][:, :10], atol=1e-10)


def test_sequential_pop_control_path(golden):
    """walkers.overlap_energy=False: comb first, energy afterwards on the copied walkers (the
    reference's literal order); must give the same trace as the overlapped default."""
    g = golden('stress_comb')
    afqmc, h = _run(g, g['h1e'], g['hs_pot'], float(g['ecore']), walkers={'overlap_energy': False})
    assert numpy.array_equal(h['parent_ix'], g['parent_ix'])
    _close(h['unscaled_weight'], g['unscaled_weight'], atol=1e-13)
    _close(h['ot'], g['ot'])
    _close(h['eloc'], g['eloc'], atol=1e-10)
    _close(afqmc.estimators.rows()[:, :10], g['rows'][:, :10], atol=1e-10)


def test_walker_restart_round_trip(golden, tmp_path):
    """write_walkers / read_walkers (walkers/handler.py:432-485): a run continued from the restart
    record of another run reproduces that run's walkers, overlaps and weights."""
    g = golden('stress_comb')
    base = str(tmp_path / 'restart.h5')
    nelec = tuple(int(x) for x in g['nelec'])
    system = Generic(nelec=nelec, h1e=numpy.array([g['h1e'], g['h1e']]), chol=g['hs_pot'],
                     ecore=float(g['ecore']))
    opts = _options(g, walkers={'write_freq': 7, 'write_file': base})
    a = AFQMC(options=opts, system=system, verbose=0)
    a.run(verbose=0)                    # 30 steps: restart written at 7, 14, 21, 28
    rec = numpy.load(base + '.rank0.npy')
    assert rec.shape == (int(g['nwalkers']), 3 + g['h1e'].shape[0] * sum(nelec))
    b = AFQMC(options=_options(g, walkers={'read_file': base}), system=system, verbose=0)
    numpy.testing.assert_array_equal(b.psi.get_write_buffers(), rec)
    # the record is the state after step 28 of the first run
    a2 = AFQMC(options=opts, system=system, verbose=0)
    seen = {}
    a2.run(verbose=0, observer=lambda step, q: seen.setdefault(step, q.psi.get_write_buffers())
           if step == 28 else None)
    numpy.testing.assert_array_equal(seen[28], rec)
    out = a.estimators.dump_datasets(str(tmp_path / 'estimates.0.npz'), metadata={'qmc': opts['qmc']})
    z = numpy.load(out)
    assert [h.decode() for h in z['basic/headers']][:3] == ['WeightFactor', 'Weight', 'ENumer']
    numpy.testing.assert_array_equal(z['basic/energies/%09d' % 2], a.estimators.rows()[2][1:])


def test_mixed_one_rdm(golden):
    """Mixed one-body density matrix (estimators/mixed.py:226-229,279-283) against the reference."""
    g = golden('mixed_rdm')
    nelec = tuple(int(x) for x in g['nelec'])
    system = Generic(nelec=nelec, h1e=numpy.array([g['h1e'], g['h1e']]), chol=g['hs_pot'],
                     ecore=float(g['ecore']))
    opts = _options(g)
    opts['estimates']['mixed']['one_rdm'] = True
    afqmc = AFQMC(options=opts, system=system, verbose=0)
    afqmc.run(verbose=0)
    _close(afqmc.estimators.rows()[:, :10], g['rows'][:, :10], atol=1e-10)
    rdm = numpy.array(afqmc.estimators.estimators['mixed'].one_rdm)
    _close(rdm, g['mixed_one_rdm'].real, rtol=1e-9, atol=1e-10)


def test_reference_driver_goldens(golden):
    """The reference's own assertions (pauxy/qmc/tests/test_afqmc.py:227,229)."""
    g = golden('test_generic')
    afqmc, _ = _run(g, g['h1e'], g['hs_pot'], float(g['ecore']))
    afqmc.estimators.estimators['mixed'].update(afqmc.system, afqmc.qmc, afqmc.trial, afqmc.psi, 0)
    numer = afqmc.estimators.estimators['mixed'].estimates[2]
    assert numer.real == pytest.approx(3.8763193646854273, rel=1e-9)
    rows = afqmc.estimators.rows()
    assert numpy.mean(rows[:-1, 5].real) == pytest.approx(1.5485077038208, rel=1e-9)


def test_pair_branch_matches_reference(golden):
    g = golden('stress_pair_branch')
    afqmc, h = _run(g, g['h1e'], g['hs_pot'], float(g['ecore']),
                    walkers={'population_control': 'pair_branch',
                             'min_weight': float(g['min_weight']),
                             'max_weight': float(g['max_weight'])})
    _close(h['weight'], g['weight'], atol=1e-13)
    _close(h['ot'], g['ot'])
    _close(h['eloc'], g['eloc'], atol=1e-10)
    _close(afqmc.estimators.rows()[:, :10], g['rows'][:, :10], atol=1e-10)


@pytest.mark.parametrize('name', ['free_comb', 'free_pair_branch', 'phaseless_nofb'])
def test_free_projection_and_no_force_bias(golden, name):
    """propagate_walker_free (continuous.py:175-200), the free-projection branches of
    orthogonalise (handler.py:178-181) and Mixed.update (mixed.py:151-177), and the phaseless
    walk with force_bias=False (continuous.py:136-138), against traces of the reference."""
    g = golden(name)
    walkers = None
    if str(g['pop_control']) == 'pair_branch':
        walkers = {'population_control': 'pair_branch', 'min_weight': float(g['min_weight']),
                   'max_weight': float(g['max_weight'])}
    afqmc, h = _run(g, g['h1e'], g['hs_pot'], float(g['ecore']), walkers=walkers,
                    propagator={'free_projection': bool(g['free_projection']),
                                'force_bias': bool(g['force_bias'])})
    if walkers is None:
        assert numpy.array_equal(h['parent_ix'], g['parent_ix'])
    _close(h['weight'], g['weight'], atol=1e-13)
    _close(h['unscaled_weight'], g['unscaled_weight'], atol=1e-13)
    _close(h['ot'], g['ot'])
    _close(h['phase'], g['phase'], atol=1e-12)
    _close(h['hybrid_energy'], g['hybrid_energy'], rtol=RTOL, atol=EH_ATOL / float(g['dt']))
    _close(h['eloc'], g['eloc'], atol=1e-10)
    assert afqmc.propagators.nhe_trig == int(g['nhe_trig'])
    assert afqmc.propagators.nfb_trig == 0
    _close(afqmc.estimators.rows()[:, :10], g['rows'][:, :10], atol=1e-10)
    _close(afqmc.psi.phi_host(), g['phi_final'], atol=1e-11)


@pytest.mark.parametrize('name', ['bp_ref', 'bp_stress', 'bp_restore_full', 'bp_restore_partial'])
def test_back_propagation(golden, name):
    """Back-propagated one-body density matrices (estimators/back_propagation.py:127-225,
    propagation/generic.py:180-213,253-290, walkers/stack.py:5-127) against traces of the
    reference; bp_ref is the reference's own driver test (qmc/tests/test_afqmc.py:232-278),
    bp_stress re-orthogonalises inside the back propagation, splits it in two and moves field
    histories between walkers in the comb."""
    g = golden(name)
    bpo = {'tau_bp': float(g['tau_bp']), 'nsplit': int(g['nsplit']), 'one_rdm': True}
    if 'restore_weights' in g and str(g['restore_weights']) != 'None':
        bpo['restore_weights'] = str(g['restore_weights'])     # back_propagation.py:75-80,187-196
    afqmc, h = _run(g, g['h1e'], g['hs_pot'], float(g['ecore']), back_propagated=bpo)
    assert numpy.array_equal(h['parent_ix'], g['parent_ix'])
    _close(h['weight'], g['weight'], atol=1e-13)
    _close(h['ot'], g['ot'], rtol=2e-9 if name == 'bp_stress' else 1e-10)   # stress walk amplifies rounding
    bp = afqmc.estimators.estimators['back_prop']
    ix = list(g['bp_buff_ix'])
    got_den = numpy.zeros(len(ix), dtype=numpy.complex128)
    got_rdm = numpy.zeros((len(ix),) + g['bp_one_rdm'].shape[1:], dtype=numpy.complex128)
    seen = {}
    for n, b in enumerate(ix):
        k = seen.get(b, 0)
        got_den[n] = bp.output['denominator'][b][k]
        got_rdm[n] = bp.output['one_rdm'][b][k]
        seen[b] = k + 1
    assert all(len(bp.output['denominator'][b]) == seen[b] for b in seen)
    _close(got_den, g['bp_denominator'], rtol=1e-12 if 'restore' not in name else 1e-10)
    _close(got_rdm, g['bp_one_rdm'], rtol=1e-9, atol=1e-9)
    _close(afqmc.engine.get_phi_bp(historic=True).cpu().numpy(), g['phi_old_final'], atol=1e-9)
    if name == 'bp_ref':
        rdm = bp.one_rdm()
        assert rdm[0, 0].trace().real == pytest.approx(3.0)
        assert rdm[0, 1].trace().real == pytest.approx(3.0)
        assert rdm[11, 0, 1, 3].real == pytest.approx(-0.121883381144845, rel=1e-8)


@pytest.mark.parametrize('name', ['c2_shape', 'c3_shape', 'c4_shape', 'c5_shape',
                                  'c2_stress', 'c3_stress', 'c4_stress', 'c5_stress'])
def test_shape_fixture_matches_reference(golden, name):
    """BASELINE shapes: `*_shape` the benchmark's benign Hamiltonian, `*_stress` scaled so that
    comb events, the force-bias clip, the hybrid-energy bound (eshift != 0 from block 2 on), the
    weight cap and several re-orthogonalisations happen AT those shapes (c4: 32 walkers x 20
    steps, 4 re-orthogonalisations)."""
    g = golden(name)
    kw = dict(scale=float(g['scale']), ramp=float(g['ramp'])) if 'scale' in g else {}
    h1e, hs, ecore = synthetic_cholesky_hamiltonian(int(g['nbasis']), int(g['nchol']),
                                                    int(g['seed_h']), **kw)
    assert h1e.sum() == g['h1e_checksum'] and hs.sum() == g['hs_checksum']
    afqmc, h = _run(g, h1e, hs, ecore)
    assert numpy.array_equal(h['parent_ix'], g['parent_ix'])
    _close(h['weight'], g['weight'], atol=1e-13)
    _close(h['ot'], g['ot'])
    _close(h['hybrid_energy'], g['hybrid_energy'], rtol=RTOL, atol=EH_ATOL / float(g['dt']))
    _close(h['eloc'], g['eloc'], atol=1e-10)
    assert afqmc.propagators.nfb_trig == int(g['nfb_trig'])
    assert afqmc.propagators.nhe_trig == int(g['nhe_trig'])
    _close(afqmc.estimators.rows()[:, :10], g['rows'][:, :10], atol=1e-10)
    _close(afqmc.psi.phi_host()[:2], g['phi_final_head'], atol=1e-11)


# ---------------------------------------------------------------------------------------------
# multi-determinant trials and the local-energy weight update (row A8')
# ---------------------------------------------------------------------------------------------
def _md_trial(g):
    from pauxy_b200.trial import MultiSlater

    def factory(system):
        if 'orbitals' in g:      # non-orthogonal expansion given by its (possibly complex) orbitals
            return MultiSlater(system, (g['coeffs'], g['orbitals']), init=g['init'])
        return MultiSlater(system, (g['coeffs'], g['occa'], g['occb']), init=g['init'])
    return factory


@pytest.mark.parametrize('name', ['md_hybrid', 'md_local_energy'])
def test_multi_det_walker_reference_tests(golden, name):
    """The reference's own multi-determinant propagation tests on the device
    (pauxy/propagation/tests/test_generic.py:52-70 local-energy weight update, :72-92 hybrid): one
    walker, ten propagation steps at the complex eshift = trial.energy; known final weights
    0.68797524675701 / 0.7430443466368197."""
    import torch
    g = golden(name)
    nelec = tuple(int(x) for x in g['nelec'])
    system = Generic(nelec=nelec, h1e=numpy.array([g['h1e'], g['h1e']]), chol=g['hs_pot'], ecore=0.0)
    trial = _md_trial(g)(system)
    opts = {'qmc': {'timestep': 0.005, 'steps': 10, 'blocks': 1, 'rng_seed': 7, 'num_walkers': 1,
                    'stabilise_freq': 5},
            'propagator': {'hybrid': bool(g['hybrid'])},
            'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}}}
    afqmc = AFQMC(options=opts, system=system, trial=trial, verbose=0)
    eng = afqmc.engine
    assert abs(trial.energy - g['trial_energy'][0]) < 1e-12 * abs(trial.energy)
    assert abs(eng.ot[0].item() - g['init_ot']) <= 1e-11 * abs(g['init_ot'])
    for s in range(10):
        eng.propagate(g['xi'][s][None, :].copy(), eshift=complex(g['trial_energy'][0]), step=1)
        w, ot = eng.weight[0].item(), eng.ot[0].item()
        assert w == pytest.approx(float(g['weight'][s]), rel=1e-10)
        assert abs(ot - g['ot'][s]) <= 1e-10 * abs(g['ot'][s])
        numpy.testing.assert_allclose(eng.ovlp_det[:, 0].cpu().numpy(), g['ovlps'][s], rtol=1e-10)
        if bool(g['hybrid']):
            assert abs(eng.hybrid_energy[0].item() - g['hybrid_energy'][s]) <= 1e-9 * abs(g['hybrid_energy'][s])
        else:
            assert abs(eng.walker_eloc[0].item() - g['eloc'][s]) <= 1e-10 * abs(g['eloc'][s])
    assert eng.weight[0].item() == pytest.approx(float(g['ref_test_golden_weight']), rel=1e-9)


@pytest.mark.parametrize('name', ['md_driver', 'md_driver_le', 'md_nomsd_cplx'])
def test_multi_det_driver(golden, name):
    """Whole driver loop with a 3-determinant particle-hole trial (MultiDetWalker population, comb,
    re-orthogonalisation, local_energy_multi_det in the mixed estimator) against traces of the
    reference; md_driver_le uses the local-energy weight update (propagator.hybrid = false)."""
    g = golden(name)
    afqmc, h = _run(g, g['h1e'], g['hs_pot'], float(g['ecore']),
                    propagator={'hybrid': bool(g['hybrid'])}, trial_factory=_md_trial(g))
    assert afqmc.psi.walker_type == 'MSD'
    assert numpy.array_equal(h['parent_ix'], g['parent_ix'])
    _close(h['weight'], g['weight'], atol=1e-13)
    _close(h['unscaled_weight'], g['unscaled_weight'], atol=1e-13)
    _close(h['ot'], g['ot'])
    _close(h['eloc'], g['eloc'], atol=1e-10)
    if bool(g['hybrid']):
        _close(h['hybrid_energy'], g['hybrid_energy'], rtol=RTOL, atol=EH_ATOL / float(g['dt']))
    assert afqmc.propagators.nfb_trig == int(g['nfb_trig'])
    assert afqmc.propagators.nhe_trig == int(g['nhe_trig'])
    _close(afqmc.estimators.rows()[:, :10], g['rows'][:, :10], atol=1e-10)
    _close(afqmc.psi.phi_host(), g['phi_final'], atol=1e-11)

"""Host logic, C-ABI surface and multi-rank plumbing (no GPU needed)."""
import ctypes
import os
import re
import sys

import numpy
import pytest
import torch

from oracle import afqmc_oracle as orc
from helpers import host_setup
from pauxy_b200 import _lib
from pauxy_b200.walkers import pair_branch_plan, plan_moves

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from pauxy_b200 import build
    build.build()
    header = open(os.path.join(ROOT, 'include', 'pauxy_b200.h')).read()
    declared = set(re.findall(r'\b(pxb_[a-z0-9_]+)\s*\(', header))
    assert len(declared) >= 30
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), "missing export %s" % name
    assert declared == set(_lib.declared_symbols())
    assert _lib.load().pxb_abi_version() == _lib.ABI_VERSION


def test_create_validates_arguments_without_gpu():
    lib = _lib.load()
    h = ctypes.c_void_p()
    bad = _lib.PxbConfig(10, 11, 3, 5, 4, 6, 0, 0, 0.005)   # nup > nbasis
    assert lib.pxb_create(ctypes.byref(h), ctypes.byref(bad)) == -1
    ok = _lib.PxbConfig(108, 21, 21, 500, 8192, 6, 0, 0, 0.005)
    assert lib.pxb_create(ctypes.byref(h), ctypes.byref(ok)) == 0
    n = ctypes.c_size_t()
    assert lib.pxb_arena_bytes(h, ctypes.byref(n)) == 0
    assert 2e9 < n.value < 8e9        # c4 working set: a few GB of the 180 GB
    # compute entry points refuse to run before the arena / hamiltonian are set
    assert lib.pxb_propagate(h, None, 0, 0, 0.0, 1, None) == -3
    assert lib.pxb_local_energy(h, None) == -3
    lib.pxb_destroy(h)
    # complex Cholesky vectors (PXB_FLAG_COMPLEX_CHOLESKY): stacked operands make the arena larger; they
    # are refused together with the Cholesky-form exchange and with back propagation
    cc = _lib.PxbConfig(108, 21, 21, 500, 8192, 6, 0, 0, 0.005, _lib.EXCHANGE_MODES['auto'],
                        _lib.FLAG_COMPLEX_CHOLESKY, 0, 1)
    assert lib.pxb_create(ctypes.byref(h), ctypes.byref(cc)) == 0
    n2 = ctypes.c_size_t()
    assert lib.pxb_arena_bytes(h, ctypes.byref(n2)) == 0 and n2.value > n.value
    lib.pxb_destroy(h)
    for mode, nbp in ((_lib.EXCHANGE_MODES['cholesky'], 0), (_lib.EXCHANGE_MODES['auto'], 5)):
        bad = _lib.PxbConfig(108, 21, 21, 500, 64, 6, 0, 0, 0.005, mode, _lib.FLAG_COMPLEX_CHOLESKY, nbp, 1)
        assert lib.pxb_create(ctypes.byref(h), ctypes.byref(bad)) == -1


def test_comb_host_bit_exact(golden):
    from pauxy_b200.engine import comb_plan_host
    rs = numpy.random.RandomState(5)
    for n in (2, 7, 32, 1000):
        for trial in range(5):
            w = rs.rand(n) ** 3
            w[rs.randint(n)] = 0.0
            w = w / (w.sum() / n)
            r = rs.rand()
            assert numpy.array_equal(comb_plan_host(w, r), orc.comb_parents(w, r, n))
    g = golden('stress_comb')
    for s in range(g['weight_prop'].shape[0]):
        wp = numpy.abs(g['weight_prop'][s])
        gw = wp / (sum(wp) / len(wp))
        assert numpy.array_equal(comb_plan_host(gw, float(g['comb_r'][s])), g['parent_ix'][s])


def test_pair_branch_plan_matches_oracle():
    rs = numpy.random.RandomState(9)
    for trial in range(20):
        w = numpy.abs(1.0 + 0.4 * rs.normal(size=24))
        draws = list(rs.rand(24))
        a = pair_branch_plan(w, iter(draws).__next__, 0.7, 1.3)
        b = orc.pair_branch_plan(w, iter(draws).__next__, 0.7, 1.3)
        assert numpy.array_equal(a[0], b[0]) and a[1] == b[1]


@pytest.mark.parametrize('name', ['test_generic', 'c1', 'stress_comb', 'cplx_driver', 'cplx_stress'])
def test_host_setup_matches_reference(golden, name):
    g = golden(name)
    system, trial, prop = host_setup(g['h1e'], g['hs_pot'], float(g['ecore']),
                                     tuple(int(x) for x in g['nelec']), float(g['dt']))
    numpy.testing.assert_allclose(prop.mf_shift, g['setup_mf_shift'], rtol=1e-12, atol=1e-14)
    numpy.testing.assert_allclose(prop.BH1, g['setup_BH1'], rtol=1e-12, atol=1e-14)
    numpy.testing.assert_allclose(trial._rchol, g['setup_rchol'], rtol=1e-12, atol=1e-14)
    numpy.testing.assert_allclose(system.h1e_mod, g['setup_h1e_mod'], rtol=1e-12, atol=1e-14)
    numpy.testing.assert_allclose(prop.mf_core, g['setup_mf_core'], rtol=1e-12)
    # half-rotated one-body matrix reproduces sum(H1*G) for the trial's own G
    h1rot = trial.half_rotated_h1(system)
    gh = numpy.concatenate(trial.GH)
    e1 = numpy.sum(system.H1[0] * trial.G[0]) + numpy.sum(system.H1[1] * trial.G[1])
    assert abs(numpy.sum(h1rot * gh) - e1) < 1e-12 * abs(e1)


def test_qmc_options_aliases():
    from pauxy_b200.qmc import QMCOpts
    q = QMCOpts({'dt': 0.01, 'nsteps': 5, 'blocks': 3, 'nwalkers': 7, 'reortho': 2,
                 'pop_control': 4, 'seed': 3}, None)
    assert (q.dt, q.nsteps, q.nblocks, q.nwalkers, q.nstblz, q.npop_control, q.rng_seed) == \
        (0.01, 5, 3, 7, 2, 4, 3)
    assert q.total_steps == 15 and q.neqlb == 200
    d = QMCOpts({}, None)
    assert (d.nwalkers, d.dt, d.nsteps, d.nblocks, d.nstblz, d.npop_control) == \
        (10, 0.005, 10, 1000, 10, 1)


def test_unsupported_modes_fail_loudly():
    from pauxy_b200.propagation import Continuous
    g_h1e = numpy.eye(4)
    hs = numpy.zeros((16, 3))
    system, trial, prop = host_setup(g_h1e, hs, 0.0, (1, 1), 0.01)

    class Q(object):
        dt = 0.01
        nstblz = 10
    for opts in ({'optimised': False}, {'stochastic_ri': True}):
        with pytest.raises(NotImplementedError):
            Continuous(system, trial, Q(), options=opts)
    with pytest.raises(ValueError):     # the local-energy weight update has no free-projection form
        Continuous(system, trial, Q(), options={'hybrid': False, 'free_projection': True})
    assert Continuous(system, trial, Q(), options={'hybrid': False}).hybrid is False
    # continuous.py:30-33: free projection switches the force bias off
    assert Continuous(system, trial, Q(), options={'free_projection': True}).force_bias is False
    # alternative energy evaluators (systems/generic.py:77-124): exact_eri is honoured (ERI form of
    # the exchange), the sampled / truncated ones are refused
    from pauxy_b200.systems import Generic
    assert Generic(nelec=(1, 1), h1e=g_h1e, chol=hs, ecore=0.0, exact_eri=True).exact_eri
    for kw in ('stochastic_ri', 'pno', 'control_variate'):
        with pytest.raises(NotImplementedError):
            Generic(nelec=(1, 1), h1e=g_h1e, chol=hs, ecore=0.0, **{kw: True})


def test_plan_moves_partitions_pairs():
    pairs = [(1, 6), (9, 2), (5, 4), (12, 13), (3, 15)]
    nw = 4
    seen = []
    for rank in range(4):
        local, out, inc = plan_moves(pairs, nw, rank)
        for c, k in local:
            seen.append((c + rank * nw, k + rank * nw))
        for peer, slots in out.items():
            peer_inc = plan_moves(pairs, nw, peer)[2][rank]
            assert len(peer_inc) == len(slots)
            for s, d in zip(slots, peer_inc):
                seen.append((s + rank * nw, d + peer * nw))
    assert sorted(seen) == sorted(pairs)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pauxy_b200.comm import TorchComm
    comm = TorchComm()
    comm.warmup(torch.device('cpu'))
    nw = 4
    w = torch.arange(nw, dtype=torch.float64) + 10 * rank
    gw = comm.allgather_tensor(w)
    est = torch.full((10,), complex(rank + 1, 0.5), dtype=torch.complex128)
    comm.allreduce_sum_(est)
    # walker moves: global (clone, kill) pairs crossing ranks in both directions
    pairs = [(1, 6), (5, 2), (3, 0), (7, 4)]
    local, out, inc = plan_moves(pairs, nw, rank)
    payload = torch.arange(nw, dtype=torch.float64).reshape(nw, 1) * 100 + rank * 1000 + \
        torch.arange(3, dtype=torch.float64)
    state = payload.clone()
    for s, d in local:
        state[d] = payload[s]
    sends = [(p, payload[torch.tensor(sl)].reshape(-1).clone()) for p, sl in sorted(out.items())]
    recvs = [(p, torch.empty(len(sl) * 3, dtype=torch.float64)) for p, sl in sorted(inc.items())]
    comm.exchange(sends, recvs)
    for (p, buf), (_, sl) in zip(recvs, sorted(inc.items())):
        state[torch.tensor(sl)] = buf.reshape(len(sl), 3)
    q.put((rank, gw.tolist(), est[0].item(), state.tolist()))
    dist.destroy_process_group()


def test_two_rank_plumbing_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(60)
    for rank, gw, est, state in res:
        assert gw == [0.0, 1.0, 2.0, 3.0, 10.0, 11.0, 12.0, 13.0]
        assert est == complex(3.0, 1.0)
    # global payload after the moves: walker k holds walker c's payload
    allstate = res[0][3] + res[1][3]
    orig = [[w * 100.0 + r * 1000 + j for j in range(3)] for r in range(2) for w in range(4)]
    expect = [list(x) for x in orig]
    for c, k in [(1, 6), (5, 2), (3, 0), (7, 4)]:
        expect[k] = orig[c]
    assert allstate == expect


# --------------------------------------------------------------------------- estimator host logic
class _FakeEngine(object):
    """CPU stand-in for pauxy_b200.engine.Engine: just the buffers and call log the estimator
    classes touch (the arithmetic of the real one is covered by the GPU tests)."""

    def __init__(self, M, ne):
        self.estimates = torch.zeros(10, dtype=torch.complex128)
        self.theta_sum = torch.zeros((ne, M), dtype=torch.complex128)
        self.bp_rdm = torch.zeros((2, M, M), dtype=torch.complex128)
        self.bp_denom = torch.zeros(1, dtype=torch.complex128)
        self.calls = []
        self.steps = 0

    def bp_steps(self):
        return self.steps

    def back_propagate(self, n, nstblz, init_walker=False):
        self.calls.append(('bp', n, nstblz, init_walker))
        self.bp_rdm += float(n)
        self.bp_denom += 2.0

    def bp_reset(self):
        self.calls.append(('reset',))
        self.steps = 0

    def bp_zero(self):
        self.bp_rdm.zero_()
        self.bp_denom.zero_()

    def zero_estimates(self):
        self.estimates.zero_()
        self.theta_sum.zero_()


class _Qmc(object):
    dt = 0.01
    nsteps = 5
    nstblz = 4


def test_back_propagation_estimator_schedule():
    """BackPropagation.update / print_step (estimators/back_propagation.py:127-225, :282-333):
    back propagation at every split point with the configurations stored so far, history reset
    and phi_old refresh at the last one, one output record per accumulated print."""
    from pauxy_b200.estimators import BackPropagation
    from pauxy_b200.comm import SingleComm
    system, trial, prop = host_setup(numpy.eye(4), numpy.zeros((16, 3)), 0.0, (1, 1), 0.01)
    eng = _FakeEngine(4, 2)
    bp = BackPropagation({'tau_bp': 0.06, 'nsplit': 2}, True, None, _Qmc(), system, trial, complex,
                         prop.BH1, engine=eng)
    assert bp.nmax == 6 and list(bp.splits) == [3, 6]
    comm = SingleComm()
    for step in range(1, 13):
        eng.steps += 1                       # what pxb_propagate does with nbp > 0
        bp.update(system, _Qmc(), trial, None, step)
        bp.print_step(comm, 1, step)
    assert eng.calls == [('bp', 3, 4, False), ('bp', 6, 4, False), ('reset',)] * 2
    assert sorted(bp.output['denominator']) == [3, 6]
    assert [len(v) for v in bp.output['one_rdm'].values()] == [2, 2]
    # accumulators are zeroed after every print: each record holds one back propagation only
    assert bp.output['denominator'][6] == [2.0 + 0j, 2.0 + 0j]
    assert float(bp.output['one_rdm'][3][0].real.max()) == 3.0
    numpy.testing.assert_allclose(bp.one_rdm()[0], numpy.full((2, 4, 4), 3.0))
    for bad in ({'evaluate_energy': True}, {'two_rdm': True}):
        with pytest.raises(NotImplementedError):
            BackPropagation(dict(tau_bp=0.06, **bad), True, None, _Qmc(), system, trial, complex,
                            prop.BH1, engine=eng)
    # restore_weights selects the estimator weights on the device (back_propagation.py:187-196)
    eng.bp_restore_weights = lambda mode: eng.calls.append(('bp_restore_weights', mode))
    bpw = BackPropagation(dict(tau_bp=0.06, restore_weights='full'), True, None, _Qmc(), system, trial,
                          complex, prop.BH1, engine=eng)
    assert bpw.restore_weights == 'full' and ('bp_restore_weights', 'full') in eng.calls
    with pytest.raises(ValueError):     # nmax = 6 steps cannot be split in 4
        BackPropagation(dict(tau_bp=0.06, nsplit=4), True, None, _Qmc(), system, trial, complex,
                        prop.BH1, engine=eng)


def test_mixed_print_step_block_arithmetic():
    """Mixed.print_step (estimators/mixed.py:252-289) on given accumulators: block averages,
    projected energy, shift vector, and the one-RDM normalisation of mixed.py:279-283."""
    from pauxy_b200.estimators import Mixed
    from pauxy_b200.comm import SingleComm
    from oracle import afqmc_oracle as orc
    rs = numpy.random.RandomState(3)
    M, nelec = 5, (2, 1)
    h1e = rs.normal(size=(M, M))
    h1e = h1e + h1e.T
    hs = rs.normal(size=(M * M, 4))
    system, trial, prop = host_setup(h1e, hs, 0.3, nelec, 0.01)
    eng = _FakeEngine(M, 3)
    mixed = Mixed({'energy_eval_freq': 1, 'one_rdm': True, 'verbose': False}, system, True, None,
                  _Qmc(), trial, complex, engine=eng)
    acc = rs.normal(size=10) + 1j * rs.normal(size=10)
    acc[[0, 1, 3]] = numpy.abs(acc[[0, 1, 3]]) + 5.0          # weights / denominators
    eng.estimates.copy_(torch.as_tensor(acc))
    th = rs.normal(size=(3, M)) + 1j * rs.normal(size=(3, M))
    eng.theta_sum.copy_(torch.as_tensor(th))
    mixed.print_step(SingleComm(), 1, 5)
    # the oracle's restatement of the same lines on the same accumulators
    o = orc.OracleAFQMC.__new__(orc.OracleAFQMC)
    o.nsteps, o.estimates, o.rows, o.one_rdm = 5, acc.copy(), [], False
    o.print_step(5)
    numpy.testing.assert_allclose(mixed.rows[0][1:10], o.rows[0][1:10], rtol=1e-14)
    numpy.testing.assert_allclose(mixed.eshift, o.eshift_vec, rtol=1e-14)
    gs_weight = acc[1] / 5
    G = numpy.array([trial.psi[:, :2].conj().dot(th[:2]), trial.psi[:, 2:].conj().dot(th[2:])])
    numpy.testing.assert_allclose(mixed.one_rdm[0], G.real / 5 / gs_weight, rtol=1e-13)
    assert float(eng.estimates.abs().max()) == 0.0 and float(eng.theta_sum.abs().max()) == 0.0
    with pytest.raises(NotImplementedError):
        Mixed({'energy_eval_freq': 5, 'one_rdm': True}, system, True, None, _Qmc(), trial, complex,
              engine=eng)


def test_multi_det_trial_setup_matches_reference(golden):
    """Host setup for a particle-hole multi-determinant trial: orbitals, variational energy by the
    Slater-Condon rules (multi_slater.py:153-176, estimators/mixed.py:537-572), mean-field shift
    through contract_one_body (propagation/generic.py:82-86, multi_slater.py:235-259) and the
    one-body propagator, against arrays recorded from the reference."""
    from pauxy_b200.systems import Generic
    from pauxy_b200.trial import MultiSlater
    from pauxy_b200.propagation import GenericContinuous
    g = golden('md_hybrid')
    nelec = tuple(int(x) for x in g['nelec'])
    system = Generic(nelec=nelec, h1e=numpy.array([g['h1e'], g['h1e']]), chol=g['hs_pot'], ecore=0.0)
    trial = MultiSlater(system, (g['coeffs'], g['occa'], g['occb']), init=g['init'])
    assert trial.ndets == 3 and trial.psi.shape == (3, 10, 10)
    trial.half_rotate(system)
    trial.calculate_energy(system)
    numpy.testing.assert_allclose([trial.energy, trial.e1b, trial.e2b], g['trial_energy'], rtol=1e-12)

    class Q(object):
        dt = 0.005
        nstblz = 5
    prop = GenericContinuous(system, trial, Q())
    numpy.testing.assert_allclose(prop.mf_shift, g['mf_shift'], rtol=1e-12, atol=1e-14)
    numpy.testing.assert_allclose(prop.BH1, g['BH1'], rtol=1e-12, atol=1e-14)
    # non-orthogonal expansion (random real orbitals): the host classes against the oracle's
    # restatement of mixed.py:511-535 / multi_slater.py:244-258
    from oracle import multi_det_oracle as mdo
    rs = numpy.random.RandomState(5)
    psi = rs.rand(3, 10, 10)
    coeffs = rs.rand(3) + 1j * rs.rand(3)
    trial2 = MultiSlater(system, (coeffs, psi))
    trial2.half_rotate(system)
    trial2.calculate_energy(system)
    ham = mdo.MultiDetHamiltonian(g['h1e'], g['hs_pot'], 0.0, nelec, 0.005, coeffs, psi).setup_multi_det()
    numpy.testing.assert_allclose([trial2.energy, trial2.e1b, trial2.e2b], ham.trial_energy(), rtol=1e-10)
    prop2 = GenericContinuous(system, trial2, Q())
    numpy.testing.assert_allclose(prop2.mf_shift, ham.mf_shift, rtol=1e-10, atol=1e-13)


def test_hdf5_formats_round_trip(golden, tmp_path, monkeypatch):
    """pauxy_b200/io.py writes and reads the reference's HDF5 layouts (QMCPACK Hamiltonians and
    wavefunctions, estimator output, walker restart).  h5py is not in the image: the round trips run
    against the in-memory stand-in the reference harness uses (oracle/stubs/h5py), which checks the
    dataset names, shapes and the complex [..., 2] storage, not the HDF5 library itself."""
    import os
    import sys
    stubs = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', 'stubs')
    monkeypatch.syspath_prepend(stubs)
    sys.modules.pop('h5py', None)
    from pauxy_b200 import io
    assert io.have_h5py()
    g = golden('md_hybrid')
    M = g['h1e'].shape[0]
    # dense Hamiltonian, real and complex storage
    for real in (True, False):
        fn = str(tmp_path / ('ham_%d.h5' % real))
        io.write_qmcpack_dense(g['h1e'], g['hs_pot'], (5, 5), M, enuc=0.25, filename=fn, real_chol=real)
        h, chol, enuc, nmo, na, nb = io.read_qmcpack_hamiltonian(fn)
        assert (nmo, na, nb, enuc) == (M, 5, 5, 0.25)
        numpy.testing.assert_array_equal(numpy.real(h), g['h1e'])
        numpy.testing.assert_array_equal(numpy.real(chol), g['hs_pot'])
        system = io.system_from_file(fn)
        assert system.nbasis == M and system.nelec == (5, 5) and not numpy.iscomplexobj(system.chol_vecs)
    # particle-hole and non-orthogonal wavefunctions
    fn = str(tmp_path / 'phmsd.h5')
    io.write_qmcpack_wfn(fn, (g['coeffs'], g['occa'], g['occb']), 'uhf', (5, 5), M,
                         init=(g['init'][:, :5], g['init'][:, 5:]))
    (c, oa, ob), psi0 = io.read_qmcpack_wfn(fn, nelec=(5, 5))
    numpy.testing.assert_array_equal(c, g['coeffs'])
    numpy.testing.assert_array_equal(oa, g['occa'])
    numpy.testing.assert_array_equal(ob, g['occb'])
    numpy.testing.assert_array_equal(psi0, g['init'])
    rs = numpy.random.RandomState(3)
    psi = rs.rand(2, M, 10) + 1j * rs.rand(2, M, 10)
    fn = str(tmp_path / 'nomsd.h5')
    io.write_qmcpack_wfn(fn, (numpy.array([0.6, 0.8j]), psi), 'uhf', (5, 5), M)
    (c, p2), psi0 = io.read_qmcpack_wfn(fn)
    numpy.testing.assert_array_equal(p2, psi)
    numpy.testing.assert_array_equal(psi0, psi[0])
    # estimator rows and walker records
    fn = str(tmp_path / 'estimates.0.h5')
    rows = [numpy.arange(10) + 0.5j, numpy.arange(10) * 2.0]
    io.write_estimates(fn, ['WeightFactor', 'Weight'], rows, {'qmc': {'dt': 0.01}},
                       {'one_rdm': {6: [numpy.ones((2, 3, 3))]}})
    import h5py
    with h5py.File(fn, 'r') as fh5:
        numpy.testing.assert_array_equal(fh5['basic/energies/000000001'][:], rows[1])
        assert fh5['back_propagated/one_rdm_6/000000000'].shape == (2, 3, 3)
    fn = str(tmp_path / 'restart.h5')
    buf = rs.rand(4, 7) + 0j
    io.write_walkers_h5(fn, buf[:2], 0, create=True)
    io.write_walkers_h5(fn, buf[2:], 2, create=False)
    numpy.testing.assert_array_equal(io.read_walkers_h5(fn, 1, 3), buf[1:])
    sys.modules.pop('h5py', None)


def test_hot_kernels_are_dmma_and_tma_in_the_built_library():
    """Static check of the built library (cuobjdump -sass, no GPU needed): the GEMM-like kernels of the
    path issue FP64 tensor instructions (DMMA = mma.sync.m8n8k4.f64) and are fed by TMA bulk copies
    (UBLKCP) with mbarriers (SYNCS); the warp-level kernels (Green's function, CholeskyQR2) use DMMA
    and warp reductions.  What profiles/r02/sass_summary.txt tabulates."""
    import shutil
    import subprocess
    if shutil.which('cuobjdump') is None:
        pytest.skip('cuobjdump not available')
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    out = subprocess.run([sys.executable, os.path.join(root, 'tools', 'sass_summary.py')],
                         capture_output=True, text=True, timeout=600).stdout
    rows = {}
    for line in out.splitlines():
        if line.startswith('#') or line.startswith('kernel'):
            continue
        f = line.split()
        name = ' '.join(f[:-16])
        rows[name] = [int(x) for x in f[-16:]]      # instr regs stack DMMA UBLKCP SYNCS LDGSTS ...
    def col(name, idx):
        hits = [v for k, v in rows.items() if k.startswith(name)]
        assert hits, name
        return [h[idx] for h in hits]
    for k in ('taylor3_kernel', 'taylor2_kernel', 'exx_eri_kernel', 'gemm_tma_kernel'):
        assert min(col(k, 3)) > 0 and min(col(k, 4)) > 0 and min(col(k, 5)) > 0, k   # DMMA, UBLKCP, SYNCS
    for k in ('theta_kernel', 'cholqr_kernel'):
        assert min(col(k, 3)) > 0, k
    assert max(col('theta_kernel<3>', 13)) > 0      # REDUX: the register-resident pivot search

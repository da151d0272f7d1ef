"""Size-independent properties at BASELINE's full sizes, device RNG and
population control on the GPU."""
import numpy
import pytest
import torch

from oracle import afqmc_oracle as orc
from helpers import host_setup, make_engine, oracle_ham, random_walkers, relerr
from pauxy_b200.hamiltonians import make_config_hamiltonian, synthetic_cholesky_hamiltonian

pytestmark = pytest.mark.gpu


def _engine(name, W, total=None, nbp=0):
    h1e, hs, ecore, nelec = make_config_hamiltonian(name)
    system, trial, prop = host_setup(h1e, hs, ecore, nelec, 0.005)
    ham = oracle_ham(h1e, hs, ecore, nelec, 0.005)
    return make_engine(system, trial, prop, W, 0.005, total_walkers=total, nbp=nbp), ham


@pytest.mark.parametrize('cfg,W,R', [('c4', 8192, 8), ('c5', 2048, 4)])
def test_full_size_replicated_walkers(cfg, W, R):
    """BASELINE full sizes (c4: 8192 walkers; c5: the 2048-walker share of one of 8 GPUs) = R
    distinct walkers tiled: every copy must reproduce its representative (same arithmetic whatever
    the tile position), and the representatives must match the oracle (propagation + local energy)."""
    eng, ham = _engine(cfg, W)
    base = random_walkers(ham, R, seed=3)
    rs = numpy.random.RandomState(4)
    xib = rs.normal(size=(R, ham.nchol))
    reps = numpy.arange(W) % R
    eng.set_phi(torch.as_tensor(base).to(eng.device)[torch.as_tensor(reps).to(eng.device)].contiguous())
    eng.ot.copy_(torch.as_tensor(orc.calc_overlap(ham, base)[reps]))
    xi = torch.as_tensor(xib).to(eng.device)[torch.as_tensor(reps).to(eng.device)].contiguous()
    eng.propagate(xi, eshift=0.0, step=1)
    eng.local_energy()
    eng.synchronize()
    phi = eng.get_phi().cpu().numpy()
    eloc = eng.eloc.cpu().numpy()
    wt = eng.weight.cpu().numpy()
    # copies identical to their representative
    assert numpy.array_equal(phi, phi[reps])
    assert numpy.array_equal(eloc, eloc[reps])
    assert numpy.array_equal(wt, wt[reps])
    # representatives against the oracle
    tha, thb, ovlp_old = orc.greens_function(ham, base)
    p1 = orc.kinetic_real(ham, base)
    xbar, _ = orc.force_bias(ham, tha, thb)
    x, cmf, cfb, _ = orc.shift_fields(ham, xib, xbar)
    p3 = orc.kinetic_real(ham, orc.apply_exponential(p1, orc.construct_vhs(ham, x)))
    assert relerr(phi[:R], p3) < 1e-11
    t2a, t2b, ovlp_new = orc.greens_function(ham, p3)
    assert relerr(eloc[:R], orc.local_energy(ham, t2a, t2b)) < 1e-11
    assert relerr(eng.ot.cpu().numpy()[:R], ovlp_new) < 1e-11
    for i in range(R):
        wr, _, _, _ = orc.update_weight_hybrid(ham, 1.0, complex(ovlp_old[i]), complex(ovlp_new[i]),
                                               0j, complex(cfb[i]), complex(cmf[i]), 0.0)
        assert abs(wt[i] - wr) < 1e-11 * abs(wr)


def test_philox_fields():
    """Device RNG: N(0,1) statistics, determinism, and invariance of the stream of a
    GLOBAL walker index under the split of walkers over devices."""
    W = 1024
    eng, ham = _engine('c2', W)
    N = ham.nchol
    eng.propagate(None, step=3, seed=11)
    xi_a = (eng.xshifted + eng.xbar).cpu().numpy()
    assert numpy.abs(xi_a.imag).max() < 1e-12
    xi_a = xi_a.real
    n = xi_a.size
    assert abs(xi_a.mean()) < 5.0 / numpy.sqrt(n)
    assert abs(xi_a.var() - 1.0) < 5.0 * numpy.sqrt(2.0 / n)
    assert abs((xi_a ** 4).mean() - 3.0) < 0.15
    assert abs(numpy.corrcoef(xi_a[:, 0], xi_a[:, 1])[0, 1]) < 0.2
    # same (seed, step) -> same fields; other step -> different
    eng2, _ = _engine('c2', W)
    eng2.propagate(None, step=3, seed=11)
    assert numpy.array_equal((eng2.xshifted + eng2.xbar).cpu().numpy().real, xi_a)
    eng2.propagate(None, step=4, seed=11)
    assert not numpy.array_equal((eng2.xshifted + eng2.xbar).cpu().numpy().real[:8], xi_a[:8])
    # second half of the population on "another device": same fields for the same global index
    eng3, _ = _engine('c2', W // 2, total=W)
    eng3.propagate(None, step=3, seed=11, walker_offset=W // 2)
    xi_b = (eng3.xshifted + eng3.xbar).cpu().numpy().real
    numpy.testing.assert_allclose(xi_b, xi_a[W // 2:], rtol=0, atol=1e-12)


def test_device_comb_bit_exact_large():
    """Comb over 8192 walkers on the device against the oracle's sequential restatement
    (walkers/handler.py:271-301): parent_ix bit-exact, clones copied onto kills."""
    W = 8192
    eng, ham = _engine('c1', W)
    rs = numpy.random.RandomState(8)
    phi = random_walkers(ham, W, seed=5)
    eng.set_phi(phi)
    w0 = numpy.abs(1.0 + 0.5 * rs.normal(size=W))
    w0[rs.randint(0, W, 40)] = 0.0
    eng.weight.copy_(torch.as_tensor(w0))
    marks = rs.normal(size=W) + 1j * rs.normal(size=W)
    eng.hybrid_energy.copy_(torch.as_tensor(marks))
    r = 0.6180339887
    eng.pop_control_comb(r)
    eng.synchronize()
    total = sum(w0)
    gw = w0 / (total / W)
    parents = orc.comb_parents(gw, r, W)
    assert numpy.array_equal(eng.parent_ix.cpu().numpy()[:W], parents)
    assert eng.total_weight.item() == total
    expect_unscaled = w0.copy()          # the clone's whole buffer lands on the killed walker
    for c, k in orc.comb_pairs(parents):
        expect_unscaled[k] = w0[c]
    assert numpy.array_equal(eng.unscaled_weight.cpu().numpy(), expect_unscaled)
    assert numpy.all(eng.weight.cpu().numpy() == 1.0)
    out = eng.get_phi().cpu().numpy()
    eh = eng.hybrid_energy.cpu().numpy()
    expect_phi, expect_eh = phi.copy(), marks.copy()
    for c, k in orc.comb_pairs(parents):
        expect_phi[k] = phi[c]
        expect_eh[k] = marks[c]
    assert numpy.array_equal(out, expect_phi)
    assert numpy.array_equal(eh, expect_eh)


@pytest.mark.parametrize('W,kind', [(2, 'skew'), (5, 'zeros'), (3000, 'wide'), (3000, 'zeros'),
                                    (4097, 'skew'), (8192, 'equal')])
def test_device_comb_selection_edge_cases(W, kind):
    """parent_ix and the zipped (clone, kill) list for awkward weight vectors and uniforms
    (walkers/handler.py:271-301): sizes off the block size, many dead walkers, a few walkers
    holding most of the weight, all-equal weights, r at both ends of [0, 1)."""
    eng, ham = _engine('c1', W)
    rs = numpy.random.RandomState(W + len(kind))
    for r in (0.0, 1e-300, 0.5, 0.9999999999999999, float(rs.rand())):
        if kind == 'skew':
            w0 = numpy.abs(rs.normal(size=W)) * 1e-3
            w0[rs.randint(0, W, max(1, W // 500))] = 0.09 * W
        elif kind == 'zeros':
            w0 = numpy.abs(1.0 + rs.normal(size=W))
            w0[rs.rand(W) < 0.6] = 0.0
            w0[0] = 1.0
        elif kind == 'wide':
            w0 = 10.0 ** rs.uniform(-12, 2, size=W)
        else:
            w0 = numpy.ones(W)
        eng.weight.copy_(torch.as_tensor(w0))
        eng.pop_control_comb(r)
        eng.synchronize()
        total = sum(w0)
        gw = w0 / (total / W)
        try:
            parents = orc.comb_parents(gw, r, W)
        except IndexError:      # the reference's sweep runs off the end (last tooth >= total)
            continue
        assert numpy.array_equal(eng.parent_ix.cpu().numpy()[:W], parents)
        pairs = eng.pairs.cpu().numpy()
        expect = orc.comb_pairs(parents)
        assert int(pairs[0]) == len(expect)
        assert [tuple(x) for x in pairs[1:1 + 2 * len(expect)].reshape(-1, 2)] == \
            [tuple(x) for x in expect]


@pytest.mark.parametrize('name,W', [('c2', 37), ('c4', 12)])
def test_back_propagation_at_config_shapes(name, W):
    """pxb_back_propagate at BASELINE shapes (M = 24 and M = 108: one and several output tiles,
    ragged walker count) against the oracle's restatement of propagation/generic.py:253-290 and
    estimators/back_propagation.py:150-205: back-propagated determinants and sum_w w G_w."""
    nbp, nstblz = 4, 2
    eng, ham = _engine(name, W, nbp=nbp)
    rs = numpy.random.RandomState(17)
    phi0 = random_walkers(ham, W, seed=9)
    eng.set_phi(phi0)
    eng.bp_reset()                      # phi_old = the walkers we start from
    xs = []
    for step in range(1, nbp + 1):
        eng.propagate(rs.normal(size=(W, ham.nchol)), eshift=0.0, step=step)
        xs.append(eng.xshifted.cpu().numpy().copy())
    assert eng.bp_steps() == nbp
    wts = numpy.abs(1.0 + 0.3 * rs.normal(size=W))
    wts[3] = 0.0
    eng.weight.copy_(torch.as_tensor(wts))
    eng.back_propagate(nbp, nstblz)
    got_bp = eng.get_phi_bp().cpu().numpy()
    numpy.testing.assert_array_equal(eng.get_phi_bp(historic=True).cpu().numpy(), phi0)
    xs = numpy.array(xs)                # [step, W, N]
    M, na = ham.nbasis, ham.nup
    rdm = numpy.zeros((2, M, M), dtype=numpy.complex128)
    for w in range(W):
        phi_bp = ham.psi.copy()
        orc.back_propagate(ham, phi_bp, xs[:, w], nstblz)
        assert relerr(got_bp[w], phi_bp) < 1e-11
        rdm[0] += wts[w] * orc.gab(phi_bp[:, :na], phi0[w][:, :na]).T
        rdm[1] += wts[w] * orc.gab(phi_bp[:, na:], phi0[w][:, na:]).T
    assert relerr(eng.bp_rdm.cpu().numpy(), rdm) < 1e-11
    assert abs(eng.bp_denom.cpu().numpy()[0] - wts.sum()) < 1e-12 * W
    # a second call accumulates, pxb_bp_zero clears
    eng.back_propagate(nbp, nstblz)
    assert relerr(eng.bp_rdm.cpu().numpy(), 2 * rdm) < 1e-11
    eng.bp_zero()
    assert float(eng.bp_rdm.abs().max()) == 0.0


def test_theta_travels_with_walkers():
    """After population control the stored Theta of a cloned walker must equal a fresh
    Green's function of its (copied) phi: the estimator reuses it."""
    W = 64
    eng, ham = _engine('c2', W)
    rs = numpy.random.RandomState(2)
    eng.set_phi(random_walkers(ham, W, seed=6))
    xi = rs.normal(size=(W, ham.nchol))
    eng.propagate(xi, step=1)
    w0 = numpy.abs(1.0 + 0.8 * rs.normal(size=W))
    eng.weight.copy_(torch.as_tensor(w0))
    eng.pop_control_comb(0.3)
    eng.local_energy()            # uses the Theta that travelled
    e_reuse = eng.eloc.cpu().numpy().copy()
    phi = eng.get_phi().cpu().numpy()
    assert (eng.parent_ix.cpu().numpy()[:W] != 1).any()
    tha, thb, _ = orc.greens_function(ham, phi)
    assert relerr(e_reuse, orc.local_energy(ham, tha, thb)) < 1e-11


@pytest.mark.parametrize('overlap', [True, False])
def test_vanished_population_is_flagged_not_resurrected(golden, overlap):
    """walkers/handler.py:236-241: the reference exits when the total weight drops below 1e-8.
    The device paths raise a STICKY flag instead; the comb plan, the copies and the weight reset
    become no-ops, so a dead population is never silently cloned back to weight 1."""
    from pauxy_b200.systems import Generic
    from pauxy_b200.qmc import AFQMC
    g = golden('stress_comb')
    nelec = tuple(int(x) for x in g['nelec'])
    system = Generic(nelec=nelec, h1e=numpy.array([g['h1e'], g['h1e']]), chol=g['hs_pot'],
                     ecore=float(g['ecore']))
    opts = {'qmc': {'timestep': 0.02, 'steps': 5, 'blocks': 2, 'rng_seed': 3, 'num_walkers': 16},
            'walkers': {'overlap_energy': overlap},
            'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}}}
    a = AFQMC(options=opts, system=system, verbose=0)
    a.step(1)
    a.psi.check_total_weight()            # healthy population: no complaint
    a.engine.weight.zero_()
    for step in (2, 3):                   # the flag survives later (healthy-looking) plans
        a.psi.pop_control(a.comm, overlap_energy=overlap)
        assert float(a.engine.weight.abs().max().item()) == 0.0
    with pytest.raises(RuntimeError, match="total walker weight"):
        a.psi.check_total_weight()

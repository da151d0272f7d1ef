"""Shared builders for the tests (host classes of the product + oracle)."""
import numpy

from oracle import afqmc_oracle as orc
from pauxy_b200.systems import Generic
from pauxy_b200.trial import MultiSlater, get_trial_wavefunction
from pauxy_b200.propagation import GenericContinuous


class _Q(object):
    def __init__(self, dt, nstblz=10):
        self.dt = dt
        self.nstblz = nstblz


def host_setup(h1e, hs_pot, ecore, nelec, dt, psi=None):
    """Product-side host objects -> arrays for Engine.set_hamiltonian."""
    system = Generic(nelec=nelec, h1e=numpy.array([h1e, h1e]) if numpy.ndim(h1e) == 2 else h1e,
                     chol=hs_pot, ecore=ecore)
    if psi is None:
        trial = get_trial_wavefunction(system)
    else:
        trial = MultiSlater(system, (numpy.array([1.0 + 0j]), psi))
        trial.half_rotate(system)
    prop = GenericContinuous(system, trial, _Q(dt))
    return system, trial, prop


def make_engine(system, trial, prop, nwalkers, dt, total_walkers=None, exp_order=6,
                exchange='auto', nbp=0):
    from pauxy_b200.engine import Engine
    def im(a):
        return numpy.iscomplexobj(a) and float(numpy.abs(numpy.imag(a)).max()) > 0.0
    eng = Engine(system.nbasis, system.nup, system.ndown, system.nfields, nwalkers, dt,
                 exp_order=exp_order, total_walkers=total_walkers, exchange=exchange, nbp=nbp,
                 complex_one_body=im(prop.BH1),
                 complex_cholesky=im(system.hs_pot) or im(trial._rchol) or im(trial.psi))
    eng.set_hamiltonian(system.hs_pot, trial._rchol, prop.BH1, trial.half_rotated_h1(system),
                        trial.psi, prop.mf_shift, system.ecore)
    eng.init_walkers(trial.init, total_walkers or nwalkers)
    return eng


def oracle_ham(h1e, hs_pot, ecore, nelec, dt, psi=None):
    return orc.Hamiltonian(h1e, hs_pot, ecore, nelec, dt, psi=psi)


def random_walkers(ham, W, seed, spread=0.3):
    """Non-trivial walker matrices: trial + complex noise."""
    rs = numpy.random.RandomState(seed)
    M, ne = ham.nbasis, ham.ne
    phi = numpy.array([ham.psi.copy() for _ in range(W)])
    phi = phi + spread * (rs.normal(size=(W, M, ne)) + 1j * rs.normal(size=(W, M, ne))) / numpy.sqrt(M)
    return phi


def relerr(a, b):
    a = numpy.asarray(a)
    b = numpy.asarray(b)
    den = numpy.maximum(numpy.abs(b).max(), 1e-300)
    return float(numpy.abs(a - b).max() / den)

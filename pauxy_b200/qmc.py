"""AFQMC driver: the host loop of pauxy.qmc.afqmc.AFQMC with the per-walker
Python loops replaced by one batched device call per phase.

Same constructor / run / finalise interface, option sections and aliases as
the reference (pauxy/qmc/afqmc.py:82-199 construction, :200-255 run,
pauxy/qmc/options.py:84-122 QMCOpts, pauxy/qmc/utils.py:3-16 seeding).
"""
import time

import numpy
import torch

from .comm import SingleComm
from .engine import Engine
from .estimators import Estimators, back_propagation_options
from .propagation import get_propagator_driver
from .trial import get_trial_wavefunction
from .walkers import Walkers, get_input_value


class QMCOpts(object):
    """pauxy/qmc/options.py:84-122."""

    def __init__(self, inputs, system, verbose=False):
        self.nwalkers = get_input_value(inputs, 'num_walkers', default=10, alias=['nwalkers'])
        self.dt = get_input_value(inputs, 'timestep', default=0.005, alias=['dt'])
        self.nsteps = get_input_value(inputs, 'num_steps', default=10, alias=['nsteps', 'steps'])
        self.nblocks = get_input_value(inputs, 'blocks', default=1000,
                                       alias=['num_blocks', 'nblocks'])
        self.total_steps = self.nsteps * self.nblocks
        self.nstblz = get_input_value(inputs, 'stabilise_freq', default=10,
                                      alias=['nstabilise', 'reortho'])
        self.npop_control = get_input_value(inputs, 'pop_control_freq', default=1,
                                            alias=['npop_control', 'pop_control'])
        self.eqlb_time = get_input_value(inputs, 'equilibration_time', default=2.0,
                                         alias=['tau_eqlb'])
        self.neqlb = int(self.eqlb_time / self.dt)
        self.beta = get_input_value(inputs, 'beta', default=None)
        self.rng_seed = get_input_value(inputs, 'rng_seed', default=None,
                                        alias=['random_seed', 'seed'])
        if self.beta is not None:
            raise NotImplementedError("pauxy_b200: finite-temperature AFQMC is out of scope")


def set_rng_seed(seed, comm, per_rank_streams=False):
    """pauxy/qmc/utils.py:3-16.  The reference seeds `seed + rank`; here every
    rank keeps the SAME stream by default so that an N-device run reproduces
    the one-rank run."""
    if seed is None:
        seed = int(numpy.random.randint(0, 1e8)) if comm.rank == 0 else None
        seed = comm.bcast(seed, root=0)
    if per_rank_streams:
        seed = seed + comm.rank
    numpy.random.seed(seed)
    return seed


class AFQMC(object):
    """AFQMC driver (phaseless, generic Hamiltonian, single-determinant trial)."""

    def __init__(self, comm=None, options=None, system=None, trial=None, parallel=False,
                 verbose=False, device=None):
        options = options or {}
        comm = comm if comm is not None else SingleComm()
        self.comm = comm
        self.verbosity = int(verbose) if verbose is not None else options.get('verbosity', 0)
        if comm.rank != 0:
            self.verbosity = 0
        verbose = self.verbosity > 0
        self.root = comm.rank == 0
        self.rank = comm.rank
        self._init_time = time.time()
        self.run_time = time.asctime()
        if system is None:
            # afqmc.py:120-126 get_system: a QMCPACK-format integral file (needs h5py, pauxy_b200/io.py)
            sys_opts = get_input_value(options, 'system', default={}, alias=['model'])
            integrals = sys_opts.get('integrals', None)
            if integrals is None:
                raise ValueError("pauxy_b200: pass system=Generic(...) or options['system']['integrals']")
            from . import io
            nelec = (sys_opts['nup'], sys_opts['ndown']) if 'nup' in sys_opts else None
            system = io.system_from_file(integrals, nelec)
        self.system = system
        qmc_opt = get_input_value(options, 'qmc', default={}, alias=['qmc_options'])
        self.qmc = QMCOpts(qmc_opt, self.system, verbose=self.verbosity > 1)
        prop_opt = options.get('propagator', {})
        self.qmc.rng_seed = set_rng_seed(self.qmc.rng_seed, comm,
                                         per_rank_streams=prop_opt.get('per_rank_streams', False))
        self.cplx = True
        twf_opt = get_input_value(options, 'trial', default={}, alias=['trial_wavefunction'])
        if trial is not None:
            self.trial = trial
            if self.trial._rchol is None:
                self.trial.half_rotate(self.system)
        else:
            self.trial = get_trial_wavefunction(self.system, options=twf_opt, comm=comm,
                                                verbose=verbose)
        if comm.rank == 0:
            self.trial.calculate_energy(self.system)
        comm.barrier()
        self.propagators = get_propagator_driver(self.system, self.trial, self.qmc,
                                                 options=prop_opt, verbose=verbose)
        self.propagators.rng_seed = int(self.qmc.rng_seed)
        self.tsetup = time.time() - self._init_time
        wlk_opts = get_input_value(options, 'walkers', default={}, alias=['walker', 'walker_opts'])
        est_opts = get_input_value(options, 'estimators', default={},
                                   alias=['estimates', 'estimator'])
        # walkers per rank (afqmc.py:163-176)
        self.qmc.nwalkers = int(self.qmc.nwalkers / comm.size)
        if self.qmc.nwalkers == 0:
            self.qmc.nwalkers = 1
        self.qmc.ntot_walkers = self.qmc.nwalkers * comm.size
        s = self.system
        bp_opt = back_propagation_options(est_opts)
        nbp = int(bp_opt.get('tau_bp', 0) / self.qmc.dt) if bp_opt is not None else 0
        self.engine = Engine(s.nbasis, s.nup, s.ndown, s.nfields, self.qmc.nwalkers, self.qmc.dt,
                             exp_order=self.propagators.exp_nmax, device=device,
                             total_walkers=self.qmc.ntot_walkers,
                             exchange=('eri' if getattr(s, 'exact_eri', False) else
                                       est_opts.get('mixed', {}).get('exchange', 'auto')),
                             free_projection=self.propagators.free_projection,
                             force_bias=self.propagators.force_bias, nbp=nbp,
                             ndets=self.trial.ndets,
                             local_energy_weight=not self.propagators.hybrid,
                             complex_one_body=bool(numpy.abs(numpy.imag(
                                 self.propagators.propagator.BH1)).max() > 0.0),
                             complex_cholesky=self._complex_operands())
        p = self.propagators.propagator
        t = self.trial
        self.engine.set_hamiltonian(s.hs_pot, t.rchol(0), p.BH1, t.half_rotated_h1(s, 0), t.det(0),
                                    p.mf_shift, s.ecore)
        if t.ndets > 1:     # walkers/multi_det.py: one set of rotated operands per determinant
            self.engine.set_trial_det(0, t.coeffs[0])
            for i in range(1, t.ndets):
                self.engine.set_trial_det(i, t.coeffs[i], t.rchol(i), t.half_rotated_h1(s, i), t.det(i))
        self.propagators.bind(self.engine)
        self.estimators = Estimators(est_opts, self.root, self.qmc, self.system, self.trial,
                                     self.propagators.BT_BP, verbose, engine=self.engine)
        self.psi = Walkers(self.system, self.trial, self.qmc, self.engine, walker_opts=wlk_opts,
                           verbose=verbose, comm=comm, nprop_tot=self.estimators.nprop_tot,
                           nbp=self.estimators.nbp)
        comm.warmup(self.engine.device)
        self.setup_timers()
        self.eshift = 0
        self.sync_timers = bool(options.get('sync_timers', False))
        # one library call per step (Engine.step -> pxb_step, CUDA-graph replay) whenever the step
        # is the plain single-device one; `fused_step: false` keeps the phase-by-phase calls
        self.fused_step = bool(options.get('fused_step', True))
        if not options.get('cuda_graphs', True):
            self.engine.step_graphs(False)
        if verbose:
            self.estimators.estimators['mixed'].print_header()

    def _complex_operands(self):
        """True when system.hs_pot, the half-rotated Cholesky vectors or the trial orbitals have an
        imaginary part (systems/generic.py:126, generate_hamiltonian(cplx=True))."""
        def im(a):
            return numpy.iscomplexobj(a) and float(numpy.abs(numpy.imag(a)).max()) > 0.0
        t = self.trial
        return bool(im(self.system.hs_pot) or
                    any(im(t.rchol(i)) or im(t.det(i)) for i in range(t.ndets)))

    def _tick(self):
        if self.sync_timers:
            self.engine.synchronize()
        return time.time()

    def run(self, psi=None, comm=None, verbose=True, observer=None):
        """Open-ended random walk: the loop of pauxy/qmc/afqmc.py:200-255."""
        if psi is not None:
            self.psi = psi
        comm = comm if comm is not None else self.comm
        self.setup_timers()
        mixed = self.estimators.estimators['mixed']
        eshift = 0
        # estimates for the initial distribution of walkers
        mixed.update(self.system, self.qmc, self.trial, self.psi, 0,
                     self.propagators.free_projection)
        if verbose:
            mixed.print_step(comm, comm.size, 0, 1)
        self.eshift = eshift
        for step in range(1, self.qmc.total_steps + 1):
            self.step(step, comm)
            if observer is not None:
                observer(step, self)
        self.engine.synchronize()

    def step(self, step, comm=None):
        """One pass of the loop body of pauxy/qmc/afqmc.py:223-255 for the device batch (what
        `run` iterates and what bench.py times): re-orthogonalisation every `stabilise_freq`
        steps, propagation, population control, estimator update, block output + energy shift."""
        comm = comm if comm is not None else self.comm
        mixed = self.estimators.estimators['mixed']
        start_step = self._tick()
        evaluate = step % mixed.energy_eval_freq == 0
        if (self.fused_step and comm.size == 1 and self.engine.nbp == 0 and not mixed.calc_one_rdm
                and self.psi.pcont_method == 'comb' and self.psi.overlap and not self.psi.use_log_shift
                and len(self.estimators.estimators) == 1 and (mixed.eval_energy or not evaluate)
                and self.propagators.hybrid):
            # orthogonalise + propagate + pop_control + estimators.update of the loop body below as
            # ONE device call; RNG consumption order is unchanged (fields, then the comb's uniform)
            prop, psi = self.propagators, self.psi
            pop = step % self.qmc.npop_control == 0 and psi.ntot_walkers > 1
            xi = prop.fields(psi, self.system, step, comm)
            r = numpy.random.random() if pop else 0.0
            self.engine.step(xi, eshift=self.eshift, step=step, seed=prop.rng_seed,
                             walker_offset=psi.walker_offset, comb_r=r,
                             ortho=(step % self.qmc.nstblz == 0), pop=pop, energy=evaluate)
            prop.fields_consumed(step)
            psi._phi_cache = None
            self.tprop += self._tick() - start_step
            self._end_of_step(step, comm, mixed, start_step)
            return
        if step % self.qmc.nstblz == 0:
            start = self._tick()
            self.psi.orthogonalise(self.trial, self.propagators.free_projection)
            self.tortho += self._tick() - start
        start = self._tick()
        self.propagators.propagate_walkers(self.psi, self.system, self.trial, self.eshift, step,
                                           comm=comm)
        self.tprop += self._tick() - start
        if step % self.qmc.npop_control == 0:
            start = self._tick()
            self.psi.pop_control(comm, overlap_energy=(
                mixed.eval_energy and step % mixed.energy_eval_freq == 0))
            self.tpopc += self._tick() - start
        start = self._tick()
        self.estimators.update(self.system, self.qmc, self.trial, self.psi, step,
                               self.propagators.free_projection)
        self.testim += self._tick() - start
        self._end_of_step(step, comm, mixed, start_step)

    def _end_of_step(self, step, comm, mixed, start_step):
        """Block output, restart record, vanished-population poll and energy shift
        (afqmc.py:243-254)."""
        self.estimators.print_step(comm, comm.size, step)
        if self.psi.write_restart and step % self.psi.write_freq == 0:
            self.psi.write_walkers(comm)
        if step % self.qmc.nsteps == 0:
            self.psi.check_total_weight()   # handler.py:236-241, polled once per block
        if step < self.qmc.neqlb:
            self.eshift = mixed.get_shift(self.propagators.hybrid)
        else:
            self.eshift += (mixed.get_shift() - self.eshift)
        self.tstep += self._tick() - start_step

    def finalise(self, verbose=False):
        if self.root and verbose:
            print("# End Time: {:s}".format(time.asctime()))
            print("# Running time : {:.6f} seconds".format(time.time() - self._init_time))
            print("# Timing breakdown (per processor, per block/step):")
            print("# - Setup: {:.6f} s".format(self.tsetup))
            nsteps = max(self.qmc.nsteps, 1)
            nstblz = max(nsteps // self.qmc.nstblz, 1)
            npcon = max(nsteps // self.qmc.npop_control, 1)
            print("# - Step: {:.6f} s".format(self.tstep / nsteps))
            print("# - Orthogonalisation: {:.6f} s".format(self.tortho / nstblz))
            print("# - Propagation: {:.6f} s".format(self.tprop / nsteps))
            print("# - Estimators: {:.6f} s".format(self.testim / nsteps))
            print("# - Population control: {:.6f} s".format(self.tpopc / npcon))

    def setup_timers(self):
        self.tortho = 0
        self.tprop = 0
        self.testim = 0
        self.tpopc = 0
        self.tstep = 0

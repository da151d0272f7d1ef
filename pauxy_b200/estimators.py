"""Mixed estimator with the reference's interface and output row.

Mirrors pauxy.estimators.mixed.Mixed (mixed.py:77-131 construction, :133-233
update, :235-289 print_step, :345-360 get_shift, :460-469 enum) and the
Estimators container (pauxy/estimators/handler.py:18-162) for the phaseless
single-determinant path.  Per-walker work (Green's function, local energy,
weighted sums) runs on the device; the host sees ten complex numbers per block.
HDF5 output is replaced by an in-memory list of rows + optional .npy dump:
h5py is not part of this image.
"""
import time

import numpy
import torch

HEADER = ['Iteration', 'WeightFactor', 'Weight', 'ENumer', 'EDenom', 'ETotal', 'E1Body',
          'E2Body', 'EHybrid', 'Overlap', 'Time']


class _Names(dict):
    __getattr__ = dict.get


def get_estimator_enum(thermal=False):
    keys = ['uweight', 'weight', 'enumer', 'edenom', 'eproj', 'e1b', 'e2b', 'ehyb', 'ovlp', 'time']
    return _Names((k, v) for v, k in enumerate(keys))


def format_fixed_width_floats(floats):
    return ' '.join('{: .10e}'.format(f) for f in floats)


class Mixed(object):
    def __init__(self, mixed, system, root, filename, qmc, trial, dtype, engine=None):
        mixed = mixed or {}
        self.eval_energy = mixed.get('evaluate_energy', True)
        if mixed.get('two_rdm', None) is not None:
            raise NotImplementedError("pauxy_b200: two-body RDM accumulation is not built")
        self.calc_one_rdm = mixed.get('one_rdm', False)
        self.energy_eval_freq = mixed.get('energy_eval_freq', None)
        if self.energy_eval_freq is None:
            self.energy_eval_freq = qmc.nsteps
        if self.calc_one_rdm and self.energy_eval_freq != 1:
            # the reference then accumulates the Green's function left behind by the propagator
            # (of the walker BEFORE the step); only the per-step evaluation is mirrored
            raise NotImplementedError("pauxy_b200: mixed one_rdm needs energy_eval_freq = 1")
        if self.calc_one_rdm and trial.ndets > 1:
            raise NotImplementedError("pauxy_b200: mixed one_rdm needs a single-determinant trial")
        self.psi_trial = numpy.array(trial.psi)
        self.nup = system.nup
        self.one_rdm = []
        self.verbose = mixed.get('verbose', True)
        self.nsteps = qmc.nsteps
        self.header = list(HEADER)
        self.nreg = len(self.header[1:])
        self.names = get_estimator_enum()
        self.eshift = numpy.array([0, 0])
        self.global_estimates = numpy.zeros(self.nreg, dtype=numpy.complex128)
        self.engine = engine
        self.rows = []
        self.filename = filename
        self._t0 = time.time()

    @property
    def estimates(self):
        """Current device accumulators (host copy)."""
        es = self.engine.estimates.cpu().numpy().copy()
        return es

    def update(self, system, qmc, trial, psi, step, free_projection=False):
        """mixed.py:211-225 for every walker of the device batch."""
        eng = self.engine   # free projection: PXB_FLAG_FREE_PROJECTION selects mixed.py:151-177
        evaluate = (step % self.energy_eval_freq == 0)
        if evaluate and self.eval_energy:
            eng.local_energy()
        elif evaluate:
            eng.eloc.zero_()
        eng.accumulate(with_energy=evaluate)
        if self.calc_one_rdm and not free_projection:
            eng.accumulate_theta()          # mixed.py:226-229

    def print_step(self, comm, nprocs, step, nsteps=None, free_projection=False):
        """mixed.py:252-289: block averages, reduction over ranks, eshift."""
        if step % self.nsteps != 0:
            return
        if nsteps is None:
            nsteps = self.nsteps
        ns = self.names
        dev = self.engine.estimates
        if comm is not None and comm.size > 1:
            comm.allreduce_sum_(dev)
            if self.calc_one_rdm:
                comm.allreduce_sum_(self.engine.theta_sum)
        gs = dev.cpu().numpy().copy()
        gs[ns.time] = (time.time() - self._t0) / nprocs
        gs[ns.uweight:ns.weight + 1] /= nsteps
        gs[ns.ehyb:ns.time + 1] /= nsteps
        gs[ns.eproj] = gs[ns.enumer]
        gs[ns.eproj:ns.e2b + 1] = gs[ns.eproj:ns.e2b + 1] / gs[ns.edenom]
        gs[ns.ehyb] /= gs[ns.weight]
        gs[ns.ovlp] /= gs[ns.weight]
        self.eshift = numpy.array([gs[ns.ehyb], gs[ns.eproj]])
        self.global_estimates = gs
        if comm is None or comm.rank == 0:
            if self.verbose:
                print(format_fixed_width_floats([step] + list(gs[:ns.time + 1].real)))
            self.rows.append(numpy.array([step] + list(gs[:ns.time + 1])))
            if self.calc_one_rdm:
                # sum_w w Re(G_w), G_w = conj(psi) Theta_w  ->  Re(conj(psi) sum_w w Theta_w);
                # block average and normalisation of mixed.py:279-283
                th = self.engine.theta_sum.cpu().numpy()
                na, psi = self.nup, self.psi_trial
                G = numpy.array([psi[:, :na].conj().dot(th[:na]), psi[:, na:].conj().dot(th[na:])])
                self.one_rdm.append(G.real / nsteps / gs[ns.weight])
        self.zero()

    def print_header(self, eol='', encode=False):
        print(' '.join('{:>17s}'.format(h) for h in self.header) + eol)

    def get_shift(self, hybrid=True):
        return self.eshift[0].real if hybrid else self.eshift[1].real

    def zero(self):
        self.engine.zero_estimates()
        self._t0 = time.time()


def back_propagation_options(estimates):
    """estimators/handler.py:84-86: the 'back_propagation' section, alias 'back_propagated'."""
    estimates = estimates or {}
    bp = estimates.get('back_propagation', None)
    return bp if bp is not None else estimates.get('back_propagated', None)


class BackPropagation(object):
    """pauxy.estimators.back_propagation.BackPropagation (back_propagation.py:17-125 options,
    :127-225 update_uhf, :282-333 print_step) for a single-determinant trial and the generic
    Hamiltonian: back-propagated one-body density matrix with BP-PhL weights.  The walkers' field
    histories and phi_old live on the device; one pxb_back_propagate call does every walker."""

    def __init__(self, bp, root, filename, qmc, system, trial, dtype, BT2, engine=None):
        self.tau_bp = bp.get('tau_bp', 0)
        self.nmax = int(self.tau_bp / qmc.dt)
        self.header = ['E', 'E1b', 'E2b']
        self.calc_one_rdm = bp.get('one_rdm', True)
        self.init_walker = bp.get('init_walker', False)
        self.nsplit = bp.get('nsplit', 1)
        self.splits = numpy.array([(i + 1) * (self.nmax // self.nsplit) for i in range(self.nsplit)])
        self.nreg = len(self.header)
        self.accumulated = False
        self.eval_energy = bp.get('evaluate_energy', False)
        self.restore_weights = bp.get('restore_weights', None)
        if self.restore_weights not in (None, 'full', 'partial'):
            self.restore_weights = 'partial'      # back_propagation.py:190-195: anything but "full"
        for key, off in (('two_rdm', None), ('evaluate_ekt', False), ('evaluate_energy', False)):
            if bp.get(key, off) not in (off,):
                raise NotImplementedError("pauxy_b200: back_propagated option %r is not built "
                                          "(one_rdm with BP-PhL weights is)" % key)
        if self.nmax < 1:
            raise ValueError("back_propagated: tau_bp < timestep")
        if self.nmax % self.nsplit != 0:
            # the reference's FieldConfig.reset only rewinds when step % nprop_tot == 0 (stack.py:122-125),
            # which never happens then: the history would overflow
            raise ValueError("back_propagated: int(tau_bp / timestep) must be a multiple of nsplit")
        if trial.ndets != 1:
            raise NotImplementedError("pauxy_b200: back propagation needs a single-determinant trial")
        self.nstblz = qmc.nstblz
        self.BT2 = BT2
        self.dt = qmc.dt
        self.engine = engine
        if engine is not None and self.restore_weights is not None:
            engine.bp_restore_weights(self.restore_weights)
        self.G = numpy.zeros((2, system.nbasis, system.nbasis), dtype=numpy.complex128)
        self.buff_ix = 0
        # what the reference pushes to estimates.h5 under back_propagated/ (in memory here)
        self.output = {'denominator': {}, 'one_rdm': {}}

    def update(self, system, qmc, trial, psi, step, free_projection=False):
        eng = self.engine
        buff_ix = eng.bp_steps()
        if buff_ix not in self.splits:
            return
        eng.back_propagate(buff_ix, self.nstblz, self.init_walker)
        if buff_ix == self.splits[-1]:
            eng.bp_reset()
        self.accumulated = True
        self.buff_ix = buff_ix

    def print_step(self, comm, nprocs, step, nsteps=1, free_projection=False):
        if not self.accumulated:
            return
        eng = self.engine
        if comm is not None and comm.size > 1:
            comm.allreduce_sum_(eng.bp_rdm)
            comm.allreduce_sum_(eng.bp_denom)
        if comm is None or comm.rank == 0:
            weight = complex(eng.bp_denom.cpu().numpy()[0])
            self.output['denominator'].setdefault(self.buff_ix, []).append(weight)
            if self.calc_one_rdm:
                self.output['one_rdm'].setdefault(self.buff_ix, []).append(eng.bp_rdm.cpu().numpy().copy())
        self.accumulated = False
        self.zero()

    def zero(self):
        self.engine.bp_zero()

    def one_rdm(self, buff_ix=None):
        """Normalised back-propagated density matrices [nprints, 2, M, M] (what
        pauxy.analysis.extraction.extract_rdm returns)."""
        ix = self.splits[-1] if buff_ix is None else buff_ix
        rdm = numpy.array(self.output['one_rdm'].get(ix, []))
        den = numpy.array(self.output['denominator'].get(ix, []))
        return rdm / den[:, None, None, None]


class Estimators(object):
    """pauxy/estimators/handler.py:18-162, mixed estimator only."""

    def __init__(self, estimates, root, qmc, system, trial, BT2, verbose=False, engine=None):
        estimates = estimates or {}
        if estimates.get('itcf') is not None:
            raise NotImplementedError("pauxy_b200: imaginary-time correlation functions are not built")
        self.basename = estimates.get('basename', 'estimates')
        self.filename = estimates.get('filename', None)
        self.estimators = {'mixed': Mixed(estimates.get('mixed', {}), system, root, self.filename,
                                          qmc, trial, complex, engine=engine)}
        bp = back_propagation_options(estimates)
        self.back_propagation = bp is not None
        if self.back_propagation:
            self.estimators['back_prop'] = BackPropagation(bp, root, self.filename, qmc, system,
                                                           trial, complex, BT2, engine=engine)
            self.nprop_tot = self.estimators['back_prop'].nmax
            self.nbp = self.estimators['back_prop'].nmax
        else:
            self.nprop_tot = None
            self.nbp = None
        self.calc_itcf = False

    def update(self, system, qmc, trial, psi, step, free_projection=False):
        for k, e in self.estimators.items():
            e.update(system, qmc, trial, psi, step, free_projection)

    def print_step(self, comm, nprocs, step, nmeasure=1, free_projection=False):
        for k, e in self.estimators.items():
            e.print_step(comm, nprocs, step, free_projection=free_projection)

    def rows(self):
        """Block rows pushed so far ([Iteration] + 10 estimates, complex)."""
        return numpy.array(self.estimators['mixed'].rows)

    def dump(self, filename=None):
        filename = filename or (self.basename + '.0.npy')
        numpy.save(filename, self.rows())
        return filename

    def dump_datasets(self, filename=None, metadata=None):
        """Everything the reference writes to estimates.<n>.h5, under the same dataset names
        (estimators/utils.py:308-320 H5EstimatorHelper.push, estimators/mixed.py:368-371,
        estimators/back_propagation.py:340-345, estimators/handler.py:119-121), in an .npz
        container: 'basic/headers', 'basic/energies/%09d', 'back_propagated/denominator_<ix>/%09d',
        'back_propagated/one_rdm_<ix>/%09d', 'metadata'."""
        import json
        from . import io
        if filename is not None and str(filename).endswith(('.h5', '.hdf5')) and io.have_h5py():
            bp = self.estimators.get('back_prop')
            io.write_estimates(filename, self.estimators['mixed'].header[1:],
                               [row[1:] for row in self.estimators['mixed'].rows], metadata,
                               bp.output if bp is not None else None)
            return filename
        filename = filename or (self.basename + '.0.npz')
        out = {'basic/headers': numpy.array(self.estimators['mixed'].header[1:]).astype('S')}
        for n, row in enumerate(self.estimators['mixed'].rows):
            out['basic/energies/%09d' % n] = numpy.asarray(row[1:])
        bp = self.estimators.get('back_prop')
        if bp is not None:
            for ix, vals in bp.output['denominator'].items():
                for n, v in enumerate(vals):
                    out['back_propagated/denominator_%d/%09d' % (ix, n)] = numpy.array([v])
            for ix, vals in bp.output['one_rdm'].items():
                for n, v in enumerate(vals):
                    out['back_propagated/one_rdm_%d/%09d' % (ix, n)] = v
        out['metadata'] = numpy.array(json.dumps(metadata or {}))
        numpy.savez(filename, **out)
        return filename

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
        if mixed.get('one_rdm', False) or mixed.get('two_rdm', None) is not None:
            raise NotImplementedError("pauxy_b200: RDM accumulation is not built")
        self.energy_eval_freq = mixed.get('energy_eval_freq', None)
        if self.energy_eval_freq is None:
            self.energy_eval_freq = qmc.nsteps
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
        self.zero()

    def print_header(self, eol='', encode=False):
        print(' '.join('{:>17s}'.format(h) for h in self.header) + eol)

    def get_shift(self, hybrid=True):
        return self.eshift[0].real if hybrid else self.eshift[1].real

    def zero(self):
        self.engine.zero_estimates()
        self._t0 = time.time()


class Estimators(object):
    """pauxy/estimators/handler.py:18-162, mixed estimator only."""

    def __init__(self, estimates, root, qmc, system, trial, BT2, verbose=False, engine=None):
        estimates = estimates or {}
        for key in ('back_propagation', 'back_propagated', 'itcf'):
            if estimates.get(key) is not None:
                raise NotImplementedError("pauxy_b200: %s is a 'next' row (SURVEY 8f.1)" % key)
        self.basename = estimates.get('basename', 'estimates')
        self.filename = estimates.get('filename', None)
        self.estimators = {'mixed': Mixed(estimates.get('mixed', {}), system, root, self.filename,
                                          qmc, trial, complex, engine=engine)}
        self.back_propagation = False
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

"""TEST / BENCH INFRASTRUCTURE -- times the CPU restatement (oracle port) of the
hot path on one host core:  python -m oracle.cpu_baseline c4 4 3
prints one JSON line {"walker_steps": n, "seconds": t}.  bench.py launches one
such process per host core (single-threaded BLAS), mirroring the reference's
one-MPI-rank-per-core execution model."""
import json
import os
import sys
import time

for _v in ('OPENBLAS_NUM_THREADS', 'OMP_NUM_THREADS', 'MKL_NUM_THREADS'):
    os.environ.setdefault(_v, '1')

import numpy  # noqa: E402

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import afqmc_oracle as orc  # noqa: E402
from pauxy_b200.hamiltonians import make_config_hamiltonian, CONFIGS  # noqa: E402


def main():
    name = sys.argv[1]
    nwalkers = int(sys.argv[2])
    nsteps = int(sys.argv[3])
    seed = int(sys.argv[4]) if len(sys.argv) > 4 else 0
    h1e, hs, ecore, nelec = make_config_hamiltonian(name)
    ham = orc.Hamiltonian(h1e, hs, ecore, nelec, 0.005)
    o = orc.OracleAFQMC(ham, nwalkers, nsteps=nsteps, nblocks=1,
                        nstblz=CONFIGS[name]['stabilise_freq'])
    numpy.random.seed(100 + seed)
    t0 = time.time()
    for _ in range(nsteps):
        xi = numpy.random.normal(0.0, 1.0, (int(o.active_mask().sum()), ham.nchol))
        o.do_step(xi, numpy.random.random)
    dt = time.time() - t0
    print(json.dumps({'walker_steps': nwalkers * nsteps, 'seconds': dt,
                      'etotal': float(numpy.array(o.rows)[-1, 5].real)}))


if __name__ == '__main__':
    main()

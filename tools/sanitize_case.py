"""Small end-to-end cases for compute-sanitizer (tools/gpu.sh sanitize):
    compute-sanitizer --tool memcheck|racecheck python tools/sanitize_case.py
Runs (1) the c1 smoke walk (32 walkers x 10 steps, comb every step, one re-orthogonalisation, local
energy, block output), (2) a c2-shaped walk with 64 walkers so that the TMA-fed persistent kernels
(taylor2, gemm_tma, exx_eri, theta) run under the tool, c3- and c4-shaped walks with a handful of
walkers for the taylor3 variants, CholeskyQR2 and the exchange item queue, (3) with >= 2 GPUs: the 64-walker stress
walk on 2 ranks with the peer-memory comb (cross-device pulls of live arenas)."""
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from pauxy_b200.hamiltonians import make_config_hamiltonian  # noqa: E402
from pauxy_b200.systems import Generic  # noqa: E402
from pauxy_b200.qmc import AFQMC  # noqa: E402


def walk(name, nwalkers, steps, blocks, stab, rng='host'):
    h1e, hs, ecore, nelec = make_config_hamiltonian(name)
    system = Generic(nelec=nelec, h1e=numpy.array([h1e, h1e]), chol=hs, ecore=ecore)
    opts = {'qmc': {'timestep': 0.005, 'steps': steps, 'blocks': blocks, 'rng_seed': 8,
                    'num_walkers': nwalkers, 'stabilise_freq': stab},
            'propagator': {'rng': rng},
            'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}}}
    a = AFQMC(options=opts, system=system, verbose=0, device=torch.device('cuda:0'))
    a.run(verbose=0)
    rows = a.estimators.rows()
    assert numpy.all(numpy.isfinite(rows.real))
    print('sanitize_case: %s %d walkers x %d steps ok, E = %.10f, launches %d' % (
        name, nwalkers, steps * blocks, rows[-1, 5].real, a.engine.launch_count()), flush=True)
    a.engine.close()


def two_rank_worker(rank, world, port):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from pauxy_b200.comm import TorchComm
    g = dict(numpy.load(os.path.join(ROOT, 'tests', 'golden', 'stress_comb64.npz')))
    nelec = tuple(int(x) for x in g['nelec'])
    system = Generic(nelec=nelec, h1e=numpy.array([g['h1e'], g['h1e']]), chol=g['hs_pot'],
                     ecore=float(g['ecore']))
    opts = {'qmc': {'timestep': float(g['dt']), 'steps': 5, 'blocks': 1, 'rng_seed': int(g['seed']),
                    'num_walkers': 64, 'stabilise_freq': 3},
            'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}}}
    comm = TorchComm()
    a = AFQMC(comm=comm, options=opts, system=system, verbose=0, device=torch.device('cuda', rank))
    a.run(comm=comm, verbose=0)
    if rank == 0:
        rows = a.estimators.rows()
        err = numpy.abs(rows[:, :10] - g['rows'][:1, :10]).max()
        print('sanitize_case: 2-rank peer comb, 64 walkers x 5 steps, peers %s, |rows - reference| %.2e'
              % (a.engine.peers_attached, err), flush=True)
    torch.cuda.synchronize()
    a.engine.close()
    dist.barrier()
    dist.destroy_process_group()


if __name__ == '__main__':
    which = sys.argv[1:] or ['c1', 'c2', 'c3', 'c4', 'multi']
    if 'c1' in which:
        walk('c1', 32, 5, 2, 5)
    if 'c2' in which:
        walk('c2', 64, 3, 1, 2, rng='philox')
    if 'c3' in which:   # taylor3 with one column group and two CTAs per SM, register-resident Gauss-Jordan / CholeskyQR2
        walk('c3', 12, 3, 1, 2, rng='philox')
    if 'c4' in which:   # taylor3 with six column groups of two warps, the dynamic exchange queue, 16 x 8 VHS tiles
        walk('c4', 8, 2, 1, 2, rng='philox')
    if 'multi' in which and torch.cuda.device_count() >= 2:
        import torch.multiprocessing as mp
        mp.start_processes(two_rank_worker, args=(2, 29733), nprocs=2, start_method='spawn')

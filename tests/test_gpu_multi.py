"""Multi-device parity: an N-device run must reproduce the ONE-rank reference run
 -- selection bit-exact, values <= 1e-10.  Needs 2 GPUs (4 / 8 for the
wider cases; the driver's test box has one, so the kept evidence is profiles/r02/
pytest_gpu_multi_n*.log and the `parity_nranks` block of every multi-GPU bench line)."""
import os

import numpy
import pytest

pytestmark = pytest.mark.gpu

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def _worker(rank, world, port, name, pop, peer, q):
    import torch
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    from pauxy_b200.comm import TorchComm
    from pauxy_b200.systems import Generic
    from pauxy_b200.qmc import AFQMC
    g = dict(numpy.load(os.path.join(GOLD, name + '.npz')))
    nelec = tuple(int(x) for x in g['nelec'])
    system = Generic(nelec=nelec, h1e=numpy.array([g['h1e'], g['h1e']]), chol=g['hs_pot'],
                     ecore=float(g['ecore']))
    opts = {'qmc': {'timestep': float(g['dt']), 'steps': int(g['steps']), 'blocks': int(g['blocks']),
                    'rng_seed': int(g['seed']), 'num_walkers': int(g['nwalkers']),
                    'stabilise_freq': int(g['stab']), 'pop_control_freq': int(g['popc'])},
            'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}}}
    opts['walkers'] = {'peer_copy': peer}
    if 'tau_bp' in g:
        opts['estimates']['back_propagated'] = {'tau_bp': float(g['tau_bp']),
                                                'nsplit': int(g['nsplit']), 'one_rdm': True}
    if pop == 'pair_branch':
        opts['walkers'].update({'population_control': 'pair_branch',
                                'min_weight': float(g['min_weight']),
                                'max_weight': float(g['max_weight'])})
    comm = TorchComm()
    afqmc = AFQMC(comm=comm, options=opts, system=system, verbose=0, device=torch.device('cuda', rank))
    hist = {k: [] for k in ('weight', 'ot', 'eloc', 'parent_ix', 'unscaled_weight')}

    def obs(step, a):
        e = a.engine
        hist['weight'].append(e.weight.cpu().numpy().copy())
        hist['unscaled_weight'].append(e.unscaled_weight.cpu().numpy().copy())
        hist['ot'].append(e.ot.cpu().numpy().copy())
        hist['eloc'].append(e.eloc.cpu().numpy().copy())
        hist['parent_ix'].append(e.parent_ix.cpu().numpy().copy())
    afqmc.run(comm=comm, verbose=0, observer=obs)
    rows = afqmc.estimators.rows() if rank == 0 else None
    bp = afqmc.estimators.estimators.get('back_prop')
    bp_out = bp.output if (bp is not None and rank == 0) else None
    q.put((rank, {k: numpy.array(v) for k, v in hist.items()}, rows,
           bool(afqmc.engine.peers_attached), bp_out))
    dist.barrier()
    dist.destroy_process_group()


def _run_n(name, pop, peer=True, world=2):
    import torch
    if torch.cuda.device_count() < world:
        pytest.skip("needs %d GPUs" % world)
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    _run_n.calls = getattr(_run_n, 'calls', 0) + 1
    port = 29600 + (os.getpid() * 7 + _run_n.calls) % 2000
    procs = [ctx.Process(target=_worker, args=(r, world, port, name, pop, peer, q))
             for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=600) for _ in procs], key=lambda x: x[0])
    for p in procs:
        p.join(120)
    return res


@pytest.mark.parametrize('world,name,pop,peer', [
    (2, 'stress_comb', 'comb', True), (2, 'c1', 'comb', True), (2, 'stress_comb', 'comb', False),
    (2, 'bp_stress', 'comb', True), (2, 'bp_stress', 'comb', False),
    (2, 'stress_pair_branch', 'pair_branch', False),
    # 64 walkers: clones cross several devices (stress_comb64: 1153 comb events in 30 steps)
    (2, 'stress_comb64', 'comb', True), (2, 'stress_pair_branch64', 'pair_branch', False),
    (4, 'stress_comb64', 'comb', True), (4, 'stress_comb64', 'comb', False),
    (4, 'stress_pair_branch64', 'pair_branch', False), (4, 'bp_stress', 'comb', True),
    (8, 'stress_comb64', 'comb', True), (8, 'stress_comb64', 'comb', False),
    (8, 'stress_pair_branch64', 'pair_branch', False)])
def test_devices_reproduce_one_rank_reference(world, name, pop, peer):
    """peer=True: clones are pulled out of the other device's arena over NVLink
    (pxb_pop_control_comb_peers); peer=False: packed NCCL send/recv planned on the host."""
    res = _run_n(name, pop, peer, world)
    if peer:
        assert all(r[3] for r in res), "CUDA IPC peer mapping of the arenas failed"
    # bp_stress is a deliberately stiff walk (Cholesky vectors x 6, dt = 0.02, 12 orbitals): it
    # amplifies rounding differences to ~4e-9 within its 20 steps on one device as well
    tol = 1e-8 if name == 'bp_stress' else 1e-10
    unscaled = numpy.concatenate([r[1]['unscaled_weight'] for r in res], axis=1)
    numpy.testing.assert_allclose(unscaled, numpy.load(os.path.join(GOLD, name + '.npz'))['unscaled_weight'],
                                  rtol=tol, atol=1e-13)
    g = dict(numpy.load(os.path.join(GOLD, name + '.npz')))
    weight = numpy.concatenate([r[1]['weight'] for r in res], axis=1)
    ot = numpy.concatenate([r[1]['ot'] for r in res], axis=1)
    eloc = numpy.concatenate([r[1]['eloc'] for r in res], axis=1)
    if pop == 'comb':
        for r in res:      # every rank computed the same, bit-exact plan
            assert numpy.array_equal(r[1]['parent_ix'], g['parent_ix'])
    numpy.testing.assert_allclose(weight, g['weight'], rtol=tol, atol=1e-13)
    numpy.testing.assert_allclose(ot, g['ot'], rtol=tol)
    numpy.testing.assert_allclose(eloc, g['eloc'], rtol=tol, atol=tol)
    numpy.testing.assert_allclose(res[0][2][:, :10], g['rows'][:, :10], rtol=tol, atol=tol)
    if 'tau_bp' in g:
        # field histories and phi_old moved between the devices with the clones: the
        # back-propagated density matrices of the 2-device run equal the 1-rank reference's
        out = res[0][4]
        seen = {}
        for n, b in enumerate(g['bp_buff_ix']):
            k = seen.get(int(b), 0)
            seen[int(b)] = k + 1
            assert abs(out['denominator'][int(b)][k] - g['bp_denominator'][n]) < 1e-10
            numpy.testing.assert_allclose(out['one_rdm'][int(b)][k], g['bp_one_rdm'][n],
                                          rtol=1e-8, atol=1e-8)

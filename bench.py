#!/usr/bin/env python
"""Benchmark of the AFQMC hot path (BASELINE.json metric: walker-steps/s
including local energy, and % of the FP64 tensor roofline).

    python bench.py --gpus 1 --steps 5 --warmup 3            # c4: M=108, 21/21, N=500, 8192 walkers
    torchrun ... bench.py --gpus N ...                        # weak scaling: 8192 walkers per GPU
    python bench.py --impl reference ...                      # CPU path (oracle port) on the host cores

One "step" = one pass of the driver loop body (pauxy/qmc/afqmc.py:223-255) over
the whole walker batch: [re-orthogonalisation every stabilise_freq steps],
propagation, comb population control, Green's function + local energy, and
the estimator accumulation.  Prints ONE JSON line on rank 0.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy  # noqa: E402


def algorithmic_flops(M, na, nb, N, order=6, exchange='eri', vhs_sym=False):
    """Algorithmic flops per walker and per CALL of each stage (SURVEY.md section 8(d)
    conventions: real x complex MAC = 4 flop, complex MAC = 8 flop, no padding counted).
    The exchange is counted in the form it is evaluated in: the symmetric quadratic form in
    the half-rotated ERI needs 2 D (D + 1) flop per spin, D = ns M; the Cholesky form
    4 N ns^2 M + 8 N ns^2."""
    ne = na + nb
    st = {
        'greens': sum(8.0 * (2 * n * n * M + 4 * n ** 3 / 3.0) for n in (na, nb)),
        'xgemm': 4.0 * N * ne * M,
        # symmetric L (real orbitals): only the upper triangle of VHS is a product, the rest a copy
        'vhs': 4.0 * N * (M * (M + 1) / 2.0 if vhs_sym else M * M),
        'one_body': 4.0 * M * M * ne,
        'taylor': 8.0 * order * M * M * ne,
        'exchange': (sum(2.0 * n * M * (n * M + 1) for n in (na, nb)) if exchange == 'eri' else
                     sum(4.0 * N * n * n * M + 8.0 * N * n * n for n in (na, nb))),
        'energy': 3 * 8.0 * N,
        'qr': sum(32.0 * (M * n * n - n ** 3 / 3.0) for n in (na, nb)),
    }
    return st


def survey_flops_per_walker_step(M, na, nb, N, order=6):
    """The SURVEY.md section 8(d) table (Cholesky-form exchange, Green's function counted twice,
    overlap separately): 267.24 MFLOP at c4.  Only used to relate walker-steps/s to the
    north-star phrasing; it is NOT what the kernels execute."""
    ne = na + nb
    g = sum(8.0 * (2 * n * n * M + 4 * n ** 3 / 3.0) for n in (na, nb))
    ov = sum(8.0 * (n * n * M + n ** 3 / 3.0) for n in (na, nb))
    return (2 * g + ov + 8.0 * M * M * ne + 8.0 * N * ne * M + 4.0 * M * M * N +
            8.0 * order * M * M * ne + 4.0 * N * (na * na + nb * nb) * M +
            8.0 * N * (na * na + nb * nb) + 8.0 * ne * M)


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        threading.Thread.__init__(self, daemon=True)
        self.index = index
        self.samples = []
        self._halt = threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q,
                                      '--format=csv,noheader,nounits'], stdout=subprocess.PIPE,
                                     stderr=subprocess.DEVNULL, text=True, timeout=5).stdout
                f = [x.strip() for x in out.strip().split(',')]
                if len(f) >= 7:
                    self.samples.append(f)
            except Exception:
                pass
            self._halt.wait(0.2)

    def stop(self):
        self._halt.set()
        self.join(6)
        sm = sorted(float(s[0]) for s in self.samples if s[0].replace('.', '').isdigit())
        names = ['hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap']
        reasons = [n for i, n in enumerate(names) if any(s[3 + i] == 'Active' for s in self.samples)]
        mx = [float(s[1]) for s in self.samples if s[1].replace('.', '').isdigit()]
        return {'sm_mhz': sm[len(sm) // 2] if sm else None, 'sm_max_mhz': mx[0] if mx else None,
                'reasons': reasons, 'samples': len(self.samples)}


def cpu_port_throughput(config, seconds_target=15.0):
    """Oracle port on every host core: P single-threaded processes, each with
    its own walkers (the reference's one-rank-per-core model).  Returns
    (walker_steps_per_s, cores, sample description)."""
    cores = os.cpu_count() or 1
    per_ws = {'c1': 3e-4, 'c2': 6e-4, 'c3': 3e-3, 'c4': 2.2e-2, 'c5': 0.3}.get(config, 2e-2)
    nsteps = 4
    nw = max(2, min(64, int(seconds_target / (per_ws * nsteps))))
    env = dict(os.environ, OPENBLAS_NUM_THREADS='1', OMP_NUM_THREADS='1', MKL_NUM_THREADS='1')
    t0 = time.time()
    procs = [subprocess.Popen([sys.executable, '-m', 'oracle.cpu_baseline', config, str(nw),
                               str(nsteps), str(i)], cwd=ROOT, env=env, stdout=subprocess.PIPE,
                              stderr=subprocess.DEVNULL, text=True) for i in range(cores)]
    ws, tmax = 0, 0.0
    for p in procs:
        out = p.communicate()[0].strip().splitlines()
        if p.returncode == 0 and out:
            r = json.loads(out[-1])
            ws += r['walker_steps']
            tmax = max(tmax, r['seconds'])
    wall = time.time() - t0
    sample = ('%d processes x %d walkers x %d steps of %s (oracle numpy port, 1 BLAS thread each; '
              'stepping time %.1f s, wall incl. setup %.1f s)' % (cores, nw, nsteps, config, tmax, wall))
    return (ws / tmax if tmax > 0 else 0.0), cores, sample


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from pauxy_b200.hamiltonians import CONFIGS
    cfg = CONFIGS[args.config]
    vals = []
    sample = ''
    cores = 1
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_port_throughput(args.config, seconds_target=8.0)
        if i >= args.warmup:
            vals.append(v)
    value = float(numpy.mean(vals)) if vals else 0.0
    wpg = args.walkers or cfg['nwalkers']
    line = {
        'impl': 'reference', 'metric': 'walker-steps/sec incl. local energy',
        'value': value, 'unit': 'walker-steps/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': None, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64 (complex128)', 'data': 'synthetic',
        'config': workload_config(args.config, cfg, wpg, args.gpus),
        'cpu_baseline': {'value': value, 'unit': 'walker-steps/s', 'cores': cores, 'kind': 'port',
                         'sample': sample},
        'e2e': {'value': value, 'unit': 'walker-steps/s', 'h2d_bytes_per_step': 0,
                'd2h_bytes_per_step': 0},
        'gpu_launches': 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(name, cfg, wpg, ngpu):
    return {'workload': '%s: synthetic Cholesky Hamiltonian nbasis=%d nocc=%d/%d nchol=%d, phaseless '
                        'AFQMC dt=0.005, RHF trial, comb every step, reortho every %d, local energy '
                        'every step' % (name, cfg['nbasis'], cfg['nelec'][0], cfg['nelec'][1],
                                        cfg['nchol'], cfg['stabilise_freq']),
            'walkers_per_gpu': wpg, 'walkers_total': wpg * ngpu,
            'l2': 'inputs larger than L2 (walker state %.0f MB per GPU)' % (
                wpg * cfg['nbasis'] * sum(cfg['nelec']) * 16 / 1e6)}


def measure_fp64_peak(torch, dev):
    n = 8192
    a = torch.randn(n, n, dtype=torch.float64, device=dev)
    b = torch.randn(n, n, dtype=torch.float64, device=dev)
    torch.matmul(a, b)
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(4):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        torch.matmul(a, b)
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return 2.0 * n ** 3 / best * 1e-9


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=6)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c4')
    ap.add_argument('--walkers', type=int, default=0, help='walkers per GPU (default: config)')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-prefetch', action='store_true',
                    help='e2e leg: copy the fields on the launch stream instead of prefetching')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)

    import torch
    from pauxy_b200.hamiltonians import CONFIGS, make_config_hamiltonian
    from pauxy_b200.systems import Generic
    from pauxy_b200.qmc import AFQMC
    from pauxy_b200.comm import SingleComm, TorchComm

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    dev = torch.device('cuda', local_rank)
    torch.cuda.set_device(dev)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group('nccl', device_id=dev)
        comm = TorchComm()
    else:
        comm = SingleComm()

    cfg = CONFIGS[args.config]
    wpg = args.walkers or cfg['nwalkers']
    M, (na, nb), N = cfg['nbasis'], cfg['nelec'], cfg['nchol']
    h1e, hs, ecore, nelec = make_config_hamiltonian(args.config)
    N = hs.shape[1]
    system = Generic(nelec=nelec, h1e=numpy.array([h1e, h1e]), chol=hs, ecore=ecore)
    opts = {'qmc': {'timestep': 0.005, 'steps': 10, 'blocks': 1000, 'rng_seed': 7,
                    'num_walkers': wpg * world, 'stabilise_freq': cfg['stabilise_freq'],
                    'pop_control_freq': 1},
            'propagator': {'rng': 'philox'},
            'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}}}
    afqmc = AFQMC(comm=comm, options=opts, system=system, verbose=0, device=dev)
    eng, psi, est, prop = afqmc.engine, afqmc.psi, afqmc.estimators, afqmc.propagators
    mixed = est.estimators['mixed']

    # fields: a few distinct pinned host buffers and device-resident copies
    rs = numpy.random.RandomState(1234 + rank)
    nbuf = 3
    xi_host = [torch.from_numpy(rs.normal(size=(wpg, N))).pin_memory() for _ in range(nbuf)]
    xi_dev = [x.to(dev) for x in xi_host]
    xi_stage = torch.empty((wpg, N), dtype=torch.float64, device=dev)
    combr = rs.rand(4096)
    res_host = torch.empty(10, dtype=torch.complex128).pin_memory()
    w_host = torch.empty(wpg, dtype=torch.float64).pin_memory()

    state = {'step': 0, 'eshift': 0.0}

    def one_step(e2e):
        state['step'] += 1
        step = state['step']
        if step % afqmc.qmc.nstblz == 0:
            psi.orthogonalise(afqmc.trial, False)
        if e2e:
            # H2D of this step's fields was started on the copy stream during the previous step
            # (Engine.prefetch_xi); the first one of a timed region is issued here
            if args.no_prefetch:
                xi_stage.copy_(xi_host[step % nbuf], non_blocking=True)
                eng.propagate(xi_stage, eshift=state['eshift'], step=step)
            else:
                xi_now = state.pop('xi_next', None)
                if xi_now is None:
                    xi_now = eng.prefetch_xi(xi_host[step % nbuf])
                eng.propagate(xi_now, eshift=state['eshift'], step=step)
                state['xi_next'] = eng.prefetch_xi(xi_host[(step + 1) % nbuf])
        else:
            eng.propagate(xi_dev[step % nbuf], eshift=state['eshift'], step=step)
        numpy.random.seed(step)         # same comb uniform on every rank
        psi.pop_control(comm, overlap_energy=True)   # what AFQMC.run does with energy_eval_freq = 1
        est.update(afqmc.system, afqmc.qmc, afqmc.trial, psi, step, False)
        if e2e:
            res_host.copy_(eng.estimates, non_blocking=True)             # D2H of the step's result
            w_host.copy_(eng.weight, non_blocking=True)
            torch.cuda.current_stream().synchronize()
        if step % 10 == 0:
            eng.zero_estimates()

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def timed(nsteps, e2e):
        sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(nsteps):
            one_step(e2e)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.time() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), 0.0)
        if e2e:
            ms = max(ms, wall)      # host-side work is part of the end-to-end path
        t = torch.tensor([ms, wall], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        state['wall_ms'] = float(t[1].item())
        return float(t[0].item())

    comm.warmup(dev)
    for _ in range(max(args.warmup, 3)):
        one_step(False)
    launches0 = eng.launch_count()
    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    # per-stage CUDA events on the launch stream, live inside the timed region
    eng.stage_times(reset=True)
    eng.profile(True)
    ms_total = timed(args.steps, False)
    wall_total = state['wall_ms']
    stage = eng.stage_times(reset=True)
    eng.profile(False)
    launches = eng.launch_count() - launches0
    for _ in range(2):          # warm the end-to-end path too (copy stream, pinned staging)
        one_step(True)
    state.pop('xi_next', None)
    ms_e2e = timed(args.steps, True)
    clocks = sampler.stop() if sampler else None

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    exchange = 'eri' if eng.exchange_is_eri() else 'cholesky'
    fl = algorithmic_flops(M, na, nb, N, afqmc.propagators.exp_nmax, exchange, eng.vhs_is_symmetric())
    peak = measure_fp64_peak(torch, dev)
    ws_total = wpg * world * args.steps
    value = ws_total / (ms_total * 1e-3)
    e2e_value = ws_total / (ms_e2e * 1e-3)
    # stage table: ms per step, launches per step, achieved TFLOP/s on the stage's algorithmic flops
    stages = {}
    step_flops = 0.0
    Np = (N + 7) // 8 * 8
    # field_kernel: reads X (both spins) and xi, writes the VHS operand x and the parity copies of
    # xbar and x; energy_kernel: reads X (both spins); 16 bytes per complex
    mem_bytes = {'field': 2 * Np * 16 + N * 8 + 3 * Np * 16, 'energy': 2 * Np * 16 + 5 * 16}
    hbm_peak = 6550.4
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    for name, (ms, calls) in stage.items():
        if calls == 0:
            continue
        per_call_ms = ms / calls
        row = {'ms_per_step': ms / args.steps, 'calls_per_step': calls / float(args.steps)}
        if name in fl:
            row['mflop_per_walker_call'] = fl[name] * 1e-6
            row['tflops'] = fl[name] * wpg / (per_call_ms * 1e-3) * 1e-12
            row['frac_of_peak'] = row['tflops'] / peak
            step_flops += fl[name] * calls / float(args.steps)
        if name in mem_bytes:
            # memory-bound stages: algorithmic bytes per walker-step against the measured HBM peak
            row['gbytes_per_s'] = mem_bytes[name] * wpg / (per_call_ms * 1e-3) * 1e-9
            row['frac_of_hbm_peak'] = row['gbytes_per_s'] / hbm_peak
        if name == 'pop_control':
            row['note'] = ('comb plan runs on a side stream beside xgemm/exchange/energy; its events '
                           'include waiting for a free SM, it is not additive to the step')
        stages[name] = row
    tensor_stages = [k for k in stages if k in ('xgemm', 'vhs', 'one_body', 'taylor', 'exchange')]
    dom = max(tensor_stages, key=lambda k: stages[k]['ms_per_step'])
    kernel_names = {'taylor': 'taylor2_kernel (exp(VHS) phi: persistent, TMA-fed DMMA, Horner)',
                    'vhs': 'gemm_tma_kernel<EpiVHS> (VHS = i sqrt(dt) L x)',
                    'taylor2': 'taylor2_kernel',
                    'exchange': ('exx_eri_kernel (Theta.K.Theta, symmetric half-rotated ERI)'
                                 if exchange == 'eri' else 'exchange_kernel (fused T = R Theta^T + trace)'),
                    'xgemm': 'gemm_tma_kernel<EpiX> (X = R^T Theta)',
                    'one_body': 'gemm_tma_kernel<EpiOF> (phi = BH1 phi)'}
    dom_ms = stage[dom][0] / stage[dom][1]
    traffic = None
    tpath = os.path.join(ROOT, 'profiles', 'traffic.json')
    if os.path.isfile(tpath):
        try:
            tj = json.load(open(tpath))
            ent = tj.get(args.config, {}).get(dom)
            if ent:       # bytes per launch measured by ncu at ent['walkers'] walkers, scaled to wpg
                traffic = ent['dram_bytes'] * (float(wpg) / ent['walkers'])
        except Exception:
            traffic = None
    survey_mflop = survey_flops_per_walker_step(M, na, nb, N, afqmc.propagators.exp_nmax) * 1e-6
    line = {
        'metric': 'walker-steps/sec incl. local energy', 'value': value, 'unit': 'walker-steps/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': max(args.warmup, 3),
        'ms_per_step': ms_total / args.steps, 'wall_ms_per_step': wall_total / args.steps,
        'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64 (complex128)', 'data': 'synthetic',
        'config': workload_config(args.config, cfg, wpg, world),
        'clocks': clocks,
        'e2e': {'value': e2e_value, 'unit': 'walker-steps/s',
                'h2d_bytes_per_step': wpg * N * 8, 'd2h_bytes_per_step': 160 + wpg * 8,
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'roofline': {'bound': 'tensor', 'kernel': kernel_names[dom],
                     'achieved': stages[dom]['tflops'], 'peak': peak, 'unit': 'TFLOP/s',
                     'frac': stages[dom]['tflops'] / peak, 'traffic': traffic,
                     'peak_source': 'cuBLAS DGEMM 8192^3 measured in this run (MEASURED_PEAKS.json '
                                    'has no FP64 entry)',
                     'kernel_ms': dom_ms,
                     'algorithmic_mflop_per_walker': fl[dom] * 1e-6,
                     'exchange_form': exchange, 'vhs_symmetric': eng.vhs_is_symmetric(),
                     'stages': stages,
                     'whole_step': {
                         'mflop_per_walker_step_executed_form': step_flops * 1e-6,
                         'tflops': step_flops * wpg / (ms_total / args.steps * 1e-3) * 1e-12,
                         'frac_of_peak': step_flops * wpg / (ms_total / args.steps * 1e-3) * 1e-12 / peak,
                         'survey_table_mflop_per_walker_step': survey_mflop,
                         'survey_equivalent_tflops': survey_mflop * 1e6 * value / world * 1e-12,
                         'note': 'survey_equivalent counts the Cholesky-form exchange of SURVEY.md '
                                 '8(d) that the ERI quadratic form does not execute'}},
    }
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample = cpu_port_throughput(args.config)
        line['cpu_baseline'] = {'value': v, 'unit': 'walker-steps/s', 'cores': cores, 'kind': 'port',
                                'sample': sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())

#!/usr/bin/env python
"""Benchmark of the AFQMC hot path (BASELINE.json metric: walker-steps/s
including local energy at 1/2/4/8 B200, and % of the FP64 tensor roofline).

    python bench.py --gpus 1 --steps 20 --warmup 5           # c4: M=108, 21/21, N=500, 8192 walkers
    torchrun ... bench.py --gpus N ...                        # weak (8192 per GPU) AND strong (8192 total)
    python bench.py --impl reference ...                      # CPU path (oracle port) on the host cores

One "step" = one call of `AFQMC.step` -- the product's own loop body, the one
`AFQMC.run` iterates (pauxy/qmc/afqmc.py:223-255) -- over the whole walker batch:
[re-orthogonalisation every stabilise_freq steps], propagation with the
auxiliary fields generated INSIDE the timed region (device Philox,
`propagator.rng = 'philox'`), comb population control, Green's function + local
energy, estimator accumulation, and every `qmc.steps` steps the block output
(reduction over ranks, device->host read of the estimates, energy shift).

The JSON line carries
  value             device-resident loop as above (fields drawn on the device)
  e2e               same loop with the fields supplied from HOST pinned memory every step
                    (H2D inside the timed region, prefetched one step ahead) and a D2H read of
                    the step's estimates + weights every step
  e2e_parity_mode   same loop with `propagator.rng = 'host'`: the reference's legacy numpy stream
                    drawn on the host inside the timed region (the bit-parity mode)
  scaling_strong    (N > 1) the BASELINE walker count split over the N devices
  other_configs     the other BASELINE shapes (N = 1: c1, c2, c3, c5 at 2048; N = 8: c5 at 16384 total)
  parity_nranks     (N > 1) a 30-step stress walk run on the N ranks before timing, compared with
                    the trace of the one-rank reference (tests/golden/stress_comb64.npz)
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


# ------------------------------------------------------------------ CPU legs
def cpu_port_throughput(config, seconds_target=15.0):
    """Oracle port on every host core: P single-threaded processes, each with
    its own walkers (the reference's one-rank-per-core model).  Returns
    (walker_steps_per_s, cores, sample description, walkers in the sample)."""
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
    sample = ('%d processes x %d walkers x %d steps of %s (oracle numpy port of the reference path, '
              '1 BLAS thread each, host-drawn fields included; stepping time %.1f s, wall incl. '
              'setup %.1f s)' % (cores, nw, nsteps, config, tmax, wall))
    return (ws / tmax if tmax > 0 else 0.0), cores, sample, cores * nw


def run_reference_arm(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return 0
    from pauxy_b200.hamiltonians import CONFIGS
    cfg = CONFIGS[args.config]
    vals = []
    sample, cores, nsample = '', 1, 0
    for i in range(args.warmup + args.steps):
        v, cores, sample, nsample = cpu_port_throughput(args.config, seconds_target=8.0)
        if i >= args.warmup:
            vals.append(v)
    value = float(numpy.mean(vals)) if vals else 0.0
    wpg = args.walkers or cfg['nwalkers']
    conf = workload_config(args.config, cfg, wpg, args.gpus)
    conf['cpu_sample'] = ('each step of this arm is a bounded sample: %d walkers in total (not %d); '
                          'CPU throughput per walker-step does not depend on the walker count'
                          % (nsample, wpg * args.gpus))
    line = {
        'impl': 'reference', 'metric': 'walker-steps/sec incl. local energy',
        'value': value, 'unit': 'walker-steps/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': None, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f64 (complex128)', 'data': 'synthetic',
        'config': conf,
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
                        'every step, block output every 10 steps' % (
                            name, cfg['nbasis'], cfg['nelec'][0], cfg['nelec'][1], cfg['nchol'],
                            cfg['stabilise_freq']),
            'walkers_per_gpu': wpg, 'walkers_total': wpg * ngpu,
            'loop': 'AFQMC.step (the loop body AFQMC.run iterates), propagator.rng=philox',
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
    del a, b
    return 2.0 * n ** 3 / best * 1e-9


# ------------------------------------------------------------------ GPU legs
class Runner(object):
    """One AFQMC driver of the product on this rank + the timing loops around AFQMC.step."""

    def __init__(self, torch, comm, dev, config, wpg, world, rng='philox', systems=None, top=None):
        from pauxy_b200.hamiltonians import CONFIGS, make_config_hamiltonian
        from pauxy_b200.systems import Generic
        from pauxy_b200.qmc import AFQMC
        self.torch, self.comm, self.dev, self.world = torch, comm, dev, world
        self.config, self.wpg = config, wpg
        cfg = CONFIGS[config]
        self.cfg = cfg
        systems = systems if systems is not None else {}
        if config not in systems:
            h1e, hs, ecore, nelec = make_config_hamiltonian(config)
            systems[config] = Generic(nelec=nelec, h1e=numpy.array([h1e, h1e]), chol=hs, ecore=ecore)
        system = systems[config]
        self.N = system.nfields
        opts = {'qmc': {'timestep': 0.005, 'steps': 10, 'blocks': 100000, 'rng_seed': 7,
                        'num_walkers': wpg * world, 'stabilise_freq': cfg['stabilise_freq'],
                        'pop_control_freq': 1},
                'propagator': {'rng': rng},
                'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}}}
        opts.update(top or {})
        self.afqmc = AFQMC(comm=comm, options=opts, system=system, verbose=0, device=dev)
        self.eng = self.afqmc.engine
        self.stepno = 0
        mixed = self.afqmc.estimators.estimators['mixed']
        mixed.update(self.afqmc.system, self.afqmc.qmc, self.afqmc.trial, self.afqmc.psi, 0, False)
        mixed.zero()
        self.res_host = torch.empty(10, dtype=torch.complex128).pin_memory()
        self.w_host = torch.empty(wpg, dtype=torch.float64).pin_memory()

    def close(self):
        # peers unmap this rank's arena (pxb_destroy closes the CUDA IPC handles) before it is freed
        self.torch.cuda.synchronize()
        self.eng.close()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
        self.afqmc = None
        self.eng = None
        self.torch.cuda.empty_cache()

    def step(self, read_back=False):
        self.stepno += 1
        self.afqmc.step(self.stepno, self.comm)
        if read_back:
            # D2H of the step's result: block accumulators + the weights
            torch = self.torch
            self.res_host.copy_(self.eng.estimates, non_blocking=True)
            self.w_host.copy_(self.eng.weight, non_blocking=True)
            torch.cuda.current_stream().synchronize()

    def sync_all(self):
        torch = self.torch
        torch.cuda.synchronize()
        if self.world > 1:
            import torch.distributed as dist
            dist.barrier()
            torch.cuda.synchronize()

    def timed(self, nsteps, read_back=False, host_work=False):
        """(ms over the nsteps steps on the device clock [max with the host clock when the leg has
        host work], host wall ms), max over ranks."""
        torch = self.torch
        self.sync_all()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.time()
        e0.record()
        for _ in range(nsteps):
            self.step(read_back)
        e1.record()
        torch.cuda.synchronize()
        wall = (time.time() - t0) * 1e3
        ms = max(e0.elapsed_time(e1), 0.0)
        if host_work:
            ms = max(ms, wall)
        t = torch.tensor([ms, wall], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            import torch.distributed as dist
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0].item()), float(t[1].item())

    def set_host_fields(self, on, nbuf=3, seed=1234):
        """Fields supplied from pinned host memory (propagators.field_source), nbuf distinct
        buffers cycled so that every step moves fresh bytes."""
        prop = self.afqmc.propagators
        if not on:
            prop.field_source = None
            prop._xi_ahead = None
            return
        torch = self.torch
        rs = numpy.random.RandomState(seed + self.comm.rank)
        bufs = [torch.from_numpy(rs.normal(size=(self.wpg, self.N))).pin_memory() for _ in range(nbuf)]
        prop.field_source = lambda step: bufs[step % nbuf]
        prop._xi_ahead = None


TOP_OPTIONS = {}


def stage_table(runner, stage, nsteps, fl, peak, hbm_peak):
    """ms per step, calls per step and achieved TFLOP/s (GB/s) of every stage from the CUDA events
    recorded on the launch stream inside the timed region."""
    wpg, N = runner.wpg, runner.N
    stages, step_flops = {}, 0.0
    Np = (N + 7) // 8 * 8
    # field_kernel: reads X (both spins) and xi, writes the VHS operand x and the parity copies of
    # xbar and x; energy_kernel: reads X (both spins); 16 bytes per complex
    mem_bytes = {'field': 2 * Np * 16 + N * 8 + 3 * Np * 16, 'energy': 2 * Np * 16 + 5 * 16}
    for name, (ms, calls) in stage.items():
        if calls == 0:
            continue
        per_call_ms = ms / calls
        row = {'ms_per_step': ms / nsteps, 'calls_per_step': calls / float(nsteps)}
        if name in fl:
            row['mflop_per_walker_call'] = fl[name] * 1e-6
            row['tflops'] = fl[name] * wpg / (per_call_ms * 1e-3) * 1e-12
            row['frac_of_peak'] = row['tflops'] / peak
            step_flops += fl[name] * calls / float(nsteps)
        if name in mem_bytes:
            row['gbytes_per_s'] = mem_bytes[name] * wpg / (per_call_ms * 1e-3) * 1e-9
            row['frac_of_hbm_peak'] = row['gbytes_per_s'] / hbm_peak
        if name == 'pop_control':
            row['note'] = ('comb plan runs on a side stream beside xgemm/exchange/energy; its events '
                           'include waiting for a free SM, it is not additive to the step')
        stages[name] = row
    return stages, step_flops


def run_config(torch, comm, dev, config, wpg, world, steps, warmup, systems, peak, hbm_peak,
               detailed=False, e2e=True, parity_mode_steps=0, profile_in_timed=False, repeats=1):
    """Times one configuration through the product loop.  Returns a dict."""
    from pauxy_b200.hamiltonians import CONFIGS
    cfg = CONFIGS[config]
    M, (na, nb) = cfg['nbasis'], cfg['nelec']
    r = Runner(torch, comm, dev, config, wpg, world, rng='philox', systems=systems, top=TOP_OPTIONS)
    eng = r.eng
    N = r.N
    comm.warmup(dev)
    # at least one re-orthogonalisation and one block output before the clock starts: the first
    # launch of a kernel loads its module lazily (0.1 - 2 ms, once per process)
    nwarm = max(warmup, 3, cfg['stabilise_freq'] + 1, 11)
    for _ in range(nwarm):
        r.step()
    out = {}
    launches0 = eng.launch_count()
    # per-stage CUDA events on the launch stream: live inside the timed region for the headline
    # config (2 x 13 events against a 20 ms step); the launch-bound small shapes are timed without
    # them and profiled in a second pass, the events would cost more than their kernels
    if detailed and profile_in_timed:
        eng.stage_times(reset=True)
        eng.profile(True)
    ms_total, wall_total = r.timed(steps)
    launches = eng.launch_count() - launches0
    for _ in range(repeats - 1):
        ms2, wall2 = r.timed(steps)
        if ms2 < ms_total:
            ms_total, wall_total = ms2, wall2
    stage = None
    if detailed:
        if not profile_in_timed:
            eng.stage_times(reset=True)
            eng.profile(True)
            r.timed(steps)
        stage = eng.stage_times(reset=True)
        eng.profile(False)
    exchange = 'eri' if eng.exchange_is_eri() else 'cholesky'
    fl = algorithmic_flops(M, na, nb, N, r.afqmc.propagators.exp_nmax, exchange, eng.vhs_is_symmetric())
    ws_total = wpg * world * steps
    nst = cfg['stabilise_freq']
    # executed-form flops per walker-step: every stage once, one-body twice, QR every nst steps
    step_mflop = (fl['greens'] + fl['xgemm'] + fl['vhs'] + 2 * fl['one_body'] + fl['taylor'] +
                  fl['exchange'] + fl['energy'] + fl['qr'] / nst) * 1e-6
    out.update({'config': config, 'warmup_done': nwarm, 'timed_steps': steps, 'walkers_per_gpu': wpg, 'walkers_total': wpg * world,
                'value': ws_total / (ms_total * 1e-3), 'ms_per_step': ms_total / steps,
                'wall_ms_per_step': wall_total / steps, 'gpu_launches': int(launches),
                'launches_per_step': launches / float(steps),
                'whole_step_frac': step_mflop * 1e6 * wpg / (ms_total / steps * 1e-3) * 1e-12 / peak,
                'mflop_per_walker_step_executed_form': step_mflop,
                'exchange_form': exchange, 'vhs_symmetric': bool(eng.vhs_is_symmetric()),
                'exp_nmax': r.afqmc.propagators.exp_nmax, 'fl': fl,
                'graph_replays': int(eng.step_graphs())})
    if detailed:
        out['stages'], _ = stage_table(r, stage, steps, fl, peak, hbm_peak)
        out['stage_raw'] = stage
    if e2e:
        r.set_host_fields(True)
        for _ in range(2):          # warm the end-to-end path too (copy stream, pinned staging)
            r.step(True)
        ms_e2e, wall_e2e = r.timed(steps, read_back=True, host_work=True)
        r.set_host_fields(False)
        out['e2e'] = {'value': ws_total / (ms_e2e * 1e-3), 'unit': 'walker-steps/s',
                      'h2d_bytes_per_step': wpg * N * 8, 'd2h_bytes_per_step': 160 + wpg * 8,
                      'ms_per_step': ms_e2e / steps,
                      'path': 'AFQMC.step with propagators.field_source = pinned host fields '
                              '(H2D prefetched one step ahead on a copy stream), D2H of the '
                              'estimates + weights and a stream synchronise every step'}
    if parity_mode_steps > 0:
        r.afqmc.propagators.rng = 'host'
        r.step()
        ms_p, wall_p = r.timed(parity_mode_steps, read_back=True, host_work=True)
        r.afqmc.propagators.rng = 'philox'
        out['e2e_parity_mode'] = {
            'value': wpg * world * parity_mode_steps / (ms_p * 1e-3), 'unit': 'walker-steps/s',
            'ms_per_step': ms_p / parity_mode_steps, 'steps': parity_mode_steps,
            'h2d_bytes_per_step': wpg * N * 8, 'd2h_bytes_per_step': 160 + wpg * 8,
            'path': "AFQMC.step with propagator.rng='host': the reference's legacy numpy stream "
                    "drawn on the host in global walker order inside the timed region (every rank "
                    "draws the whole global block); the bit-parity mode, not the production one"}
    r.close()
    return out


def multi_rank_parity(torch, comm, dev, peer_copy=True):
    """Runs the 64-walker stress walk (comb events every step, force-bias clip, hybrid-energy bound,
    weight cap, re-orthogonalisation every 3 steps) on the N ranks of this job and compares every
    step with the trace of the ONE-rank reference run (tests/golden/stress_comb64.npz, recorded from
    the unmodified reference by oracle/gen_golden.py).  Selection must be bit-exact."""
    from pauxy_b200.systems import Generic
    from pauxy_b200.qmc import AFQMC
    path = os.path.join(ROOT, 'tests', 'golden', 'stress_comb64.npz')
    g = dict(numpy.load(path))
    nelec = tuple(int(x) for x in g['nelec'])
    system = Generic(nelec=nelec, h1e=numpy.array([g['h1e'], g['h1e']]), chol=g['hs_pot'],
                     ecore=float(g['ecore']))
    opts = {'qmc': {'timestep': float(g['dt']), 'steps': int(g['steps']), 'blocks': int(g['blocks']),
                    'rng_seed': int(g['seed']), 'num_walkers': int(g['nwalkers']),
                    'stabilise_freq': int(g['stab']), 'pop_control_freq': int(g['popc'])},
            'walkers': {'peer_copy': peer_copy},
            'estimates': {'mixed': {'energy_eval_freq': 1, 'verbose': False}}}
    state = numpy.random.get_state()
    afqmc = AFQMC(comm=comm, options=opts, system=system, verbose=0, device=dev)
    hist = {k: [] for k in ('weight', 'unscaled_weight', 'ot', 'eloc', 'parent_ix')}

    def obs(step, a):
        e = a.engine
        hist['weight'].append(comm.allgather_tensor(e.weight).cpu().numpy())
        hist['unscaled_weight'].append(comm.allgather_tensor(e.unscaled_weight).cpu().numpy())
        hist['ot'].append(comm.allgather_tensor(e.ot).cpu().numpy())
        hist['eloc'].append(comm.allgather_tensor(e.eloc.reshape(-1)).cpu().numpy().reshape(-1, 3))
        hist['parent_ix'].append(e.parent_ix.cpu().numpy()[:int(g['nwalkers'])].copy())
    afqmc.run(comm=comm, verbose=0, observer=obs)
    numpy.random.set_state(state)

    def rel(a, b):
        a, b = numpy.asarray(a), numpy.asarray(b)
        return float(numpy.abs(a - b).max() / max(numpy.abs(b).max(), 1e-300))
    h = {k: numpy.array(v) for k, v in hist.items()}
    rows = afqmc.estimators.rows()
    errs = {'weight': rel(h['weight'], g['weight']),
            'unscaled_weight': rel(h['unscaled_weight'], g['unscaled_weight']),
            'ot': float(numpy.abs(h['ot'] / g['ot'] - 1.0).max()),
            'eloc': rel(h['eloc'], g['eloc'])}
    if comm.rank == 0:
        errs['rows'] = rel(rows[:, :10], g['rows'][:, :10])
    res = {'n': comm.size, 'case': 'stress_comb64 (64 walkers, 30 steps, one-rank reference trace)',
           'peer_copy': bool(peer_copy and afqmc.engine.peers_attached),
           'parent_ix_bit_exact': bool(numpy.array_equal(h['parent_ix'], g['parent_ix'])),
           'comb_events': int((g['parent_ix'] != 1).sum()),
           'cross_rank_moves': int(sum(
               ((numpy.where(p > 1)[0][:min((p > 1).sum(), (p == 0).sum())] // (len(p) // comm.size)) !=
                (numpy.where(p == 0)[0][:min((p > 1).sum(), (p == 0).sum())] // (len(p) // comm.size))).sum()
               for p in g['parent_ix'])),
           'max_rel': max(errs.values()), 'rel': errs, 'tolerance': 1e-10}
    res['ok'] = bool(res['parent_ix_bit_exact'] and res['max_rel'] <= 1e-10)
    torch.cuda.synchronize()
    afqmc.engine.close()
    if comm.size > 1:
        comm.barrier()
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=20)
    ap.add_argument('--warmup', type=int, default=5)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--config', default='c4')
    ap.add_argument('--walkers', type=int, default=0, help='walkers per GPU (default: config)')
    ap.add_argument('--scaling', default='both', choices=['weak', 'strong', 'both'],
                    help='N > 1: weak = config walkers per GPU (headline), strong = config walkers in '
                         'total; both = headline weak + scaling_strong block')
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--no-other-configs', action='store_true')
    ap.add_argument('--no-parity-check', action='store_true')
    ap.add_argument('--no-parity-mode', action='store_true')
    ap.add_argument('--no-graphs', action='store_true', help='pxb_step without CUDA-graph replay')
    ap.add_argument('--no-fused', action='store_true', help='phase-by-phase calls instead of pxb_step')
    args = ap.parse_args()
    if args.impl == 'reference':
        return run_reference_arm(args)

    import torch
    from pauxy_b200.hamiltonians import CONFIGS
    from pauxy_b200.comm import SingleComm, TorchComm
    if args.no_graphs:
        TOP_OPTIONS['cuda_graphs'] = False
    if args.no_fused:
        TOP_OPTIONS['fused_step'] = False

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
        dist = None
        comm = SingleComm()

    cfg = CONFIGS[args.config]
    wpg = args.walkers or cfg['nwalkers']
    M, (na, nb) = cfg['nbasis'], cfg['nelec']
    hbm_peak = 6550.4
    try:
        hbm_peak = float(json.load(open(os.path.join(ROOT, 'MEASURED_PEAKS.json')))['hbm_gbs'])
    except Exception:
        pass
    peak = measure_fp64_peak(torch, dev)
    systems = {}

    parity = None
    if world > 1 and not args.no_parity_check:
        parity = multi_rank_parity(torch, comm, dev)

    sampler = ClockSampler(local_rank) if rank == 0 else None
    if sampler:
        sampler.start()
    strong_only = world > 1 and args.scaling == 'strong'
    if strong_only:
        wpg = max(4, (args.walkers or cfg['nwalkers']) // world)
    main_res = run_config(torch, comm, dev, args.config, wpg, world, args.steps, args.warmup, systems,
                          peak, hbm_peak, detailed=True, e2e=True,
                          parity_mode_steps=0 if args.no_parity_mode else 3, profile_in_timed=True)
    clocks = sampler.stop() if sampler else None

    strong = None
    if world > 1 and args.scaling == 'both':
        wtot = args.walkers or cfg['nwalkers']
        wps = max(4, wtot // world)
        s = run_config(torch, comm, dev, args.config, wps, world, args.steps, args.warmup, systems,
                       peak, hbm_peak, detailed=True, e2e=False)
        strong = {'value': s['value'], 'unit': 'walker-steps/s', 'ms_per_step': s['ms_per_step'],
                  'wall_ms_per_step': s['wall_ms_per_step'], 'walkers_per_gpu': wps,
                  'walkers_total': wps * world, 'whole_step_frac': s['whole_step_frac'],
                  'launches_per_step': s['launches_per_step'],
                  'stages_ms': dict((k, round(v['ms_per_step'], 4)) for k, v in s['stages'].items()),
                  'note': 'BASELINE walker count split over the devices (strong scaling); compare with '
                          'the 1-GPU headline value of the same config'}

    others = {}
    if not args.no_other_configs and args.config == 'c4' and not args.walkers:
        todo = []
        if world == 1:
            todo = [('c1', 32), ('c2', 1024), ('c3', 4096), ('c5', 2048)]
        elif world == 8:
            todo = [('c5', 2048)]      # BASELINE c5: 16 384 walkers sharded over 8 GPUs
        for name, w in todo:
            try:
                # launch-latency-bound shapes: more steps, best of three timed regions (host jitter
                # of a fraction of a millisecond is as large as their step)
                small = name in ('c1', 'c2')
                o = run_config(torch, comm, dev, name, w, world, 100 if small else min(args.steps, 20), 3,
                               systems, peak, hbm_peak, detailed=True, e2e=False, repeats=3 if small else 1)
                others[name] = {'value': o['value'], 'ms_per_step': o['ms_per_step'],
                                'wall_ms_per_step': o['wall_ms_per_step'], 'timed_steps': o['timed_steps'],
                                'walkers_per_gpu': w, 'walkers_total': w * world,
                                'whole_step_frac': o['whole_step_frac'],
                                'launches_per_step': o['launches_per_step'],
                                'mflop_per_walker_step_executed_form': o['mflop_per_walker_step_executed_form'],
                                'stage_frac_of_peak': dict(
                                    (k, round(v['frac_of_peak'], 3)) for k, v in o['stages'].items()
                                    if 'frac_of_peak' in v and k in ('taylor', 'one_body', 'vhs', 'xgemm',
                                                                     'exchange', 'greens', 'qr')),
                                'stages_ms': dict((k, round(v['ms_per_step'], 4))
                                                  for k, v in o['stages'].items())}
            except Exception as e:     # a side measurement must not lose the headline
                others[name] = {'error': '%s: %s' % (type(e).__name__, e)}
            systems.pop(name, None)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    stages = main_res['stages']
    fl = main_res['fl']
    exchange = main_res['exchange_form']
    tensor_stages = [k for k in stages if k in ('xgemm', 'vhs', 'one_body', 'taylor', 'exchange')]
    dom = max(tensor_stages, key=lambda k: stages[k]['ms_per_step'])
    # which Taylor kernel the dispatch in csrc/pxb_api.cu (run_taylor3) picks for this shape
    nch = (na + nb + 47) // 48
    nt8 = (-(-(na + nb) // nch) + 7) // 8
    t3 = 4 <= (M + 7) // 8 <= 16 and 2 <= nt8 <= 6
    kernel_names = {'taylor': ('taylor3_kernel (exp(VHS) phi: persistent, TMA-fed DMMA, 3-product complex Horner)'
                               if t3 else 'taylor2_kernel (exp(VHS) phi: persistent, TMA-fed DMMA, Horner)'),
                    'vhs': 'gemm_tma_kernel<EpiVHS> (VHS = i sqrt(dt) L x)',
                    'exchange': ('exx_eri_kernel (Theta.K.Theta, symmetric half-rotated ERI)'
                                 if exchange == 'eri' else 'exchange_kernel (fused T = R Theta^T + trace)'),
                    'xgemm': 'gemm_tma_kernel<EpiX> (X = R^T Theta)',
                    'one_body': 'gemm_tma_kernel<EpiOF> (phi = BH1 phi)'}
    raw = main_res['stage_raw']
    dom_ms = raw[dom][0] / raw[dom][1]
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
    N = CONFIGS[args.config]['nchol']
    survey_mflop = survey_flops_per_walker_step(M, na, nb, N, main_res['exp_nmax']) * 1e-6
    value = main_res['value']
    pipe_peak = None
    try:
        pipe_peak = float(json.load(open(os.path.join(ROOT, 'profiles', 'fp64_peaks.json')))['dmma_pipe_tflops'])
    except Exception:
        pass
    line = {
        'metric': 'walker-steps/sec incl. local energy', 'value': value, 'unit': 'walker-steps/s',
        'n_gpus': world, 'steps': args.steps, 'warmup': main_res['warmup_done'],
        'ms_per_step': main_res['ms_per_step'], 'wall_ms_per_step': main_res['wall_ms_per_step'],
        'higher_is_better': True, 'scaling': 'strong' if strong_only else 'weak',
        'vs_baseline': None, 'dtype': 'f64 (complex128)', 'data': 'synthetic',
        'config': workload_config(args.config, cfg, wpg, world),
        'clocks': clocks,
        'e2e': main_res['e2e'],
        'gpu_launches': main_res['gpu_launches'],
        'cuda_graph_replays': main_res['graph_replays'],
        'roofline': {'bound': 'tensor', 'kernel': kernel_names[dom],
                     'achieved': stages[dom]['tflops'], 'peak': peak, 'unit': 'TFLOP/s',
                     'frac': stages[dom]['tflops'] / peak, 'traffic': traffic,
                     'peak_source': 'cuBLAS DGEMM 8192^3 measured in this run by bench.py (builder-'
                                    'measured: MEASURED_PEAKS.json has no FP64 entry)',
                     'peak_dmma_pipe': pipe_peak,
                     'frac_of_dmma_pipe': (stages[dom]['tflops'] / pipe_peak) if pipe_peak else None,
                     'kernel_ms': dom_ms,
                     'algorithmic_mflop_per_walker': fl[dom] * 1e-6,
                     'exchange_form': exchange, 'vhs_symmetric': main_res['vhs_symmetric'],
                     'stages': stages,
                     'whole_step': {
                         'mflop_per_walker_step_executed_form': main_res['mflop_per_walker_step_executed_form'],
                         'tflops': main_res['whole_step_frac'] * peak,
                         'frac_of_peak': main_res['whole_step_frac'],
                         'survey_table_mflop_per_walker_step': survey_mflop,
                         'survey_equivalent_tflops': survey_mflop * 1e6 * value / world * 1e-12,
                         'note': 'survey_equivalent counts the Cholesky-form exchange of SURVEY.md '
                                 '8(d) that the ERI quadratic form does not execute'}},
    }
    if 'e2e_parity_mode' in main_res:
        line['e2e_parity_mode'] = main_res['e2e_parity_mode']
    if strong is not None:
        line['scaling_strong'] = strong
    if others:
        line['other_configs'] = others
    if parity is not None:
        line['parity_nranks'] = parity
    if not args.no_cpu_baseline and world == 1:
        v, cores, sample, _ = cpu_port_throughput(args.config)
        line['cpu_baseline'] = {'value': v, 'unit': 'walker-steps/s', 'cores': cores, 'kind': 'port',
                                'sample': sample}
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == '__main__':
    sys.exit(main())

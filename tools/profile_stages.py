"""Runs each stage of the hot path once or twice on a BASELINE config (for ncu):
    ncu --set full -k regex:exchange -c 1 -o gpurun_out/ex python tools/profile_stages.py c4 2368
"""
import os
import sys

import numpy
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests'))
from helpers import host_setup, make_engine  # noqa: E402
from pauxy_b200.hamiltonians import CONFIGS, make_config_hamiltonian  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else 'c4'
    W = int(sys.argv[2]) if len(sys.argv) > 2 else CONFIGS[name]['nwalkers']
    reps = int(sys.argv[3]) if len(sys.argv) > 3 else 2
    h1e, hs, ecore, nelec = make_config_hamiltonian(name)
    system, trial, prop = host_setup(h1e, hs, ecore, nelec, 0.005)
    eng = make_engine(system, trial, prop, W, 0.005)
    xi = torch.randn(W, system.nfields, dtype=torch.float64, device=eng.device)
    times = {}

    def t(label, fn):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        times.setdefault(label, []).append(a.elapsed_time(b))

    for r in range(reps):
        t('propagate', lambda: eng.propagate(xi, eshift=0.0, step=r + 2))
        t('pop_control', lambda: eng.pop_control_comb(0.37))
        t('greens', lambda: eng.stage_greens(True))
        t('xgemm', lambda: eng.stage_force_bias_gemm())
        t('exchange', lambda: eng.stage_exchange())
        t('local_energy', lambda: eng.local_energy())
        t('orthogonalise', lambda: eng.orthogonalise())
    for k, v in times.items():
        print('%-14s %s ms' % (k, ' '.join('%9.3f' % x for x in v)))


if __name__ == '__main__':
    main()

"""Prints the interesting numbers of a bench.py JSON line:  python tools/bench_summary.py line.json"""
import json
import sys

l = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print('N=%d value %.0f ws/s  %.3f ms/step (wall %.3f)  e2e %.0f (%.3f ms)  launches/step %.1f  clocks %s' % (
    l['n_gpus'], l['value'], l['ms_per_step'], l.get('wall_ms_per_step', 0), l['e2e']['value'],
    l['e2e'].get('ms_per_step', 0), l['gpu_launches'] / float(l['steps']), l.get('clocks')))
r = l['roofline']
print('  roofline: %s  %.2f TF/s of %.2f = %.3f; whole step %.3f' % (
    r['kernel'][:40], r['achieved'], r['peak'], r['frac'], r['whole_step']['frac_of_peak']))
for k, v in r['stages'].items():
    print('   %-12s %8.3f ms  x%.1f  %s' % (k, v['ms_per_step'], v['calls_per_step'],
                                         ('%.3f of peak' % v['frac_of_peak']) if 'frac_of_peak' in v else ''))
if 'e2e_parity_mode' in l:
    print('  parity-mode e2e: %.0f ws/s (%.1f ms/step)' % (l['e2e_parity_mode']['value'],
                                                         l['e2e_parity_mode']['ms_per_step']))
if 'scaling_strong' in l:
    s = l['scaling_strong']
    print('  strong: %.0f ws/s  %.3f ms/step (wall %.3f) at %d walkers/GPU  frac %.3f  %s' % (
        s['value'], s['ms_per_step'], s['wall_ms_per_step'], s['walkers_per_gpu'], s['whole_step_frac'],
        s['stages_ms']))
for k, v in l.get('other_configs', {}).items():
    if 'error' in v:
        print('  %s: ERROR %s' % (k, v['error']))
    else:
        print('  %s: %.0f ws/s  %.4f ms/step  frac %.3f  launches/step %.1f  %s' % (
            k, v['value'], v['ms_per_step'], v['whole_step_frac'], v['launches_per_step'],
            v['stage_frac_of_peak']))
if 'parity_nranks' in l:
    print('  parity_nranks:', l['parity_nranks'])
if 'cpu_baseline' in l:
    print('  cpu: %.0f ws/s on %d cores' % (l['cpu_baseline']['value'], l['cpu_baseline']['cores']))

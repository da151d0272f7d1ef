#!/bin/bash
mkdir -p gpurun_out
for c in c5:2048 c3:4096 c2:1024; do
cfg=${c%%:*}; w=${c#*:}
echo "== bench $cfg W=$w"; timeout 300 python bench.py --config $cfg --walkers $w --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_$cfg.log 2>&1; tail -1 gpurun_out/bench_$cfg.log | python -c "
import json,sys
try:
    l=json.loads(sys.stdin.read())
    print(l['n_gpus'], l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], 'launches', l['gpu_launches'])
    for k,v in l['roofline']['stages'].items(): print('  ',k, round(v['ms_per_step'],3), round(v.get('frac_of_peak',0),3))
except Exception as e:
    print('FAILED', e)
"
tail -3 gpurun_out/bench_$cfg.log | head -2 | cut -c1-300
done

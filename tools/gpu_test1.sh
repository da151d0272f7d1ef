#!/bin/bash
# single-GPU: full GPU test suite + short bench
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['n_gpus'], l['value'], l['ms_per_step'], 'wall', l.get('wall_ms_per_step'), 'e2e', l['e2e']['value'], l['e2e']['ms_per_step'])
for k,v in l['roofline']['stages'].items(): print(k, round(v['ms_per_step'],3), round(v.get('frac_of_peak',0),3))
" || tail -30 gpurun_out/bench.log
echo "== bench no prefetch"; timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-prefetch 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['n_gpus'], l['value'], l['ms_per_step'], 'wall', l.get('wall_ms_per_step'), 'e2e', l['e2e']['value'], l['e2e']['ms_per_step'])"
echo "== bench prefetch again"; timeout 600 python bench.py --steps 12 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['n_gpus'], l['value'], l['ms_per_step'], 'wall', l.get('wall_ms_per_step'), 'e2e', l['e2e']['value'], l['e2e']['ms_per_step'])"

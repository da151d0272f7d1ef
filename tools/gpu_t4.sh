#!/bin/bash
mkdir -p gpurun_out
echo "== stage times"; timeout 120 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8 | head -3
echo "== pytest gpu"; timeout 400 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -4
echo "== bench"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'])
for k,v in l['roofline']['stages'].items(): print(k, round(v['ms_per_step'],3), round(v.get('frac_of_peak',0),3))
PY
echo "== bench c5"; timeout 300 python bench.py --config c5 --walkers 2048 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_c5.log 2>&1; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench_c5.log').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'])
for k,v in l['roofline']['stages'].items(): print(k, round(v['ms_per_step'],3), round(v.get('frac_of_peak',0),3))
PY

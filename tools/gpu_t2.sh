#!/bin/bash
mkdir -p gpurun_out
echo "== pytest stages+props"; timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_properties.py tests/test_gpu_driver.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -15
echo "== stage times (tma taylor)"; timeout 300 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8 | head -1
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'])
for k,v in l['roofline']['stages'].items(): print(k, round(v['ms_per_step'],3), round(v.get('frac_of_peak',0),3))
PY
echo "== ncu"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"taylor2_kernel" -c 1 -f -o gpurun_out/prof_t2 python tools/profile_stages.py c4 2368 1 > gpurun_out/ncu_t2.log 2>&1; tail -1 gpurun_out/ncu_t2.log

#!/bin/bash
# every command is bounded: a hung kernel must not eat the GPU budget
mkdir -p gpurun_out
echo "== stage times NG4"; timeout 120 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8 | head -1
echo "== pytest stages+props"; timeout 300 python -m pytest tests/test_gpu_stages.py tests/test_gpu_properties.py tests/test_gpu_driver.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5
echo "== stage times NG2"; PXB_TAYLOR_GROUPS=2 timeout 120 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8 | head -1
echo "== stage times NG4 nbuf1"; PXB_TAYLOR_NBUF=1 timeout 120 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8 | head -1
echo "== bench"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.log').read().strip().splitlines()[-1])
print(l['value'], l['ms_per_step'])
for k,v in l['roofline']['stages'].items(): print(k, round(v['ms_per_step'],3), round(v.get('frac_of_peak',0),3))
PY
echo "== ncu"; timeout 300 ncu --set full --clock-control none --import-source on -k regex:"taylor2_kernel" -c 1 -f -o gpurun_out/prof_t2 python tools/profile_stages.py c4 2368 1 > gpurun_out/ncu_t2.log 2>&1; tail -1 gpurun_out/ncu_t2.log

#!/bin/bash
# round-1 evidence: bench line, launch list (same command under ncu), full ncu of the hot kernels
mkdir -p gpurun_out
echo "== bench"; timeout 600 python bench.py --steps 6 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log > gpurun_out/bench.json; python - <<'PY'
import json
l=json.loads(open('gpurun_out/bench.json').read())
print(l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], 'cpu', l.get('cpu_baseline',{}).get('value'), 'clocks', l['clocks'])
PY
echo "== bench reference arm"; timeout 300 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 > gpurun_out/bench_reference.json; cut -c1-200 gpurun_out/bench_reference.json
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; wc -l gpurun_out/launches.csv
echo "== ncu full"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"exx_eri_kernel|taylor2_kernel|gemm_tma_kernel|theta_kernel" -c 12 -f -o gpurun_out/prof_hot python tools/profile_stages.py c4 2368 1 > gpurun_out/ncu_hot.log 2>&1; tail -2 gpurun_out/ncu_hot.log

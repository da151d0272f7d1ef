#!/bin/bash
mkdir -p gpurun_out
echo "== pytest bp"; timeout 900 python -m pytest tests/test_gpu_driver.py tests/test_gpu_properties.py -m gpu -q --tb=short -p no:cacheprovider -k "back_prop" 2>&1 | tail -15
for a in "" "--no-prefetch" "" "--no-prefetch"; do
echo "== bench $a"; timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline $a 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['n_gpus'], l['value'], l['ms_per_step'], 'wall', l.get('wall_ms_per_step'), 'e2e', l['e2e']['value'], l['e2e']['ms_per_step'])"
done

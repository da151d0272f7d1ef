#!/bin/bash
# A/B of an env-selected variant: tools/gpu_ab.sh VAR=value
mkdir -p gpurun_out
for v in "" "$1" "" "$1"; do
echo "== bench [$v]"; env $v timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'], ' '.join('%s=%.3f'%(k,v['ms_per_step']) for k,v in l['roofline']['stages'].items() if k in ('one_body','vhs','xgemm','exchange','taylor','greens')))"
done

#!/bin/bash
for m in x p d pd; do
echo -n "mode $m: "; BENCH_E2E_MODE=$m timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read()); print(l['value'], 'e2e', l['e2e']['value'], l['e2e']['ms_per_step'])"
done

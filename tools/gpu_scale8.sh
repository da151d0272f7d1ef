#!/bin/bash
N=${1:-8}
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; tail -1 gpurun_out/bench_n$N.log | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['n_gpus'], l['value'], l['ms_per_step'], 'wall', l.get('wall_ms_per_step'), 'e2e', l['e2e']['value'], l['e2e']['ms_per_step'])
for k,v in l['roofline']['stages'].items(): print(k, round(v['ms_per_step'],3))
" || tail -30 gpurun_out/bench_n$N.log

#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest multi"; timeout 300 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -5
echo "== bench 2 gpus"; timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2.log 2>&1; tail -1 gpurun_out/bench2.log | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['n_gpus'], l['value'], l['ms_per_step'], 'e2e', l['e2e']['value'])
for k,v in l['roofline']['stages'].items(): print(k, round(v['ms_per_step'],3))
"
echo "== bench 1 gpu"; timeout 300 python bench.py --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['n_gpus'], l['value'], l['ms_per_step'])"

#!/bin/bash
# usage: tools/gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L
nvidia-smi topo -m 2>/dev/null | head -12
if [ "$N" = "2" ]; then
echo "== pytest multi"; timeout 600 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -15
fi
echo "== bench $N gpus"; timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 8 --warmup 3 > gpurun_out/bench_n$N.log 2>&1; tail -1 gpurun_out/bench_n$N.log | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['n_gpus'], l['value'], l['ms_per_step'], 'wall', l.get('wall_ms_per_step'), 'e2e', l['e2e']['value'], l['e2e']['ms_per_step'])
for k,v in l['roofline']['stages'].items(): print(k, round(v['ms_per_step'],3))
" || tail -30 gpurun_out/bench_n$N.log
echo "== bench 1 gpu"; timeout 300 python bench.py --steps 8 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | python -c "
import json,sys
l=json.loads(sys.stdin.read())
print(l['n_gpus'], l['value'], l['ms_per_step'], 'wall', l.get('wall_ms_per_step'), 'e2e', l['e2e']['value'])"

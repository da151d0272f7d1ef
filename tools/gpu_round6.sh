#!/bin/bash
# validation of the restored tree: tests, bench (with cpu baseline), launch list, full ncu of the hot kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== stage times"; timeout 300 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8
echo "== bench"; timeout 1200 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log > gpurun_out/bench.json; cat gpurun_out/bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | tee gpurun_out/bench_reference.json
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"exchange_kernel|taylor_kernel|gemm_tma_kernel|greens_kernel" -c 8 -f -o gpurun_out/prof_hot python tools/profile_stages.py c4 2368 1 > gpurun_out/ncu_hot.log 2>&1; tail -2 gpurun_out/ncu_hot.log

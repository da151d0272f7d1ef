#!/bin/bash
# round-1 evidence, v6 tree: tests, smoke, bench (+ cpu baseline, reference arm), launch list, full ncu of the hot kernels, other configs
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
nproc
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
echo "== bench"; timeout 1200 python bench.py --steps 8 --warmup 3 > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log > gpurun_out/bench.json; cut -c1-600 gpurun_out/bench.json
echo "== bench reference arm"; timeout 600 python bench.py --impl reference --steps 1 --warmup 0 2>&1 | tail -1 | tee gpurun_out/bench_reference.json | cut -c1-400
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 260 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"exx_eri_kernel|taylor2_kernel|gemm_tma_kernel|theta_kernel" -c 11 -f -o gpurun_out/prof_hot python tools/profile_stages.py c4 2368 1 > gpurun_out/ncu_hot.log 2>&1; tail -2 gpurun_out/ncu_hot.log
bash tools/gpu_cfgs.sh

#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu (tma gemm)"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== stage times tma"; timeout 300 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8
echo "== stage times direct"; PXB_GEMM=direct timeout 300 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -3 gpurun_out/bench.log
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1

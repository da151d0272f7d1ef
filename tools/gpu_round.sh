#!/bin/bash
# One GPU visit: parity tests, FP64 peaks, smoke, bench, launch list.  Everything
# is logged under gpurun_out/ (merged back by gpurun).
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
echo "== pytest gpu"; timeout 900 python -m pytest tests -m gpu -q -x --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== pytest gpu (all, no -x)"; timeout 900 python -m pytest tests -m gpu -q --tb=line -p no:cacheprovider > gpurun_out/pytest_gpu_all.log 2>&1; tail -30 gpurun_out/pytest_gpu_all.log
echo "== fp64 peak"; timeout 120 tools/fp64_peak > gpurun_out/fp64_peak.txt 2>&1; cat gpurun_out/fp64_peak.txt
echo "== cublas"; timeout 300 python tools/cublas_fp64_peak.py 2>&1 | tail -3
echo "== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "== bench"; timeout 900 python bench.py --steps 4 --warmup 3 > gpurun_out/bench.log 2>&1; tail -5 gpurun_out/bench.log
echo "== ncu launches"; timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1; tail -2 gpurun_out/bench_ncu.log

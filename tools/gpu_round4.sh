#!/bin/bash
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q --tb=short -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -40 gpurun_out/pytest_gpu.log
echo "== stage times"; timeout 300 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -3 gpurun_out/bench.log
echo "== ncu taylor+greens"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"taylor|greens_kernel" -c 3 -f -o gpurun_out/prof_tg python tools/profile_stages.py c4 2368 1 > gpurun_out/ncu_tg.log 2>&1; tail -2 gpurun_out/ncu_tg.log

#!/bin/bash
mkdir -p gpurun_out
echo "== pytest stages"; timeout 900 python -m pytest tests/test_gpu_stages.py -m gpu -q --tb=short -p no:cacheprovider 2>&1 | tail -5
echo "== ncu launches tma"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_tma.csv python tools/profile_stages.py c4 8192 2 > gpurun_out/ps_tma.log 2>&1; tail -8 gpurun_out/ps_tma.log
echo "== ncu launches direct"; PXB_GEMM=direct timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/launches_direct.csv python tools/profile_stages.py c4 8192 2 > gpurun_out/ps_direct.log 2>&1

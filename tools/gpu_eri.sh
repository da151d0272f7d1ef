#!/bin/bash
mkdir -p gpurun_out
echo "== pytest stages"; timeout 900 python -m pytest tests/test_gpu_stages.py tests/test_gpu_properties.py -m gpu -q -x --tb=short -p no:cacheprovider 2>&1 | tail -15
echo "== stage times"; timeout 300 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench.log 2>&1; tail -1 gpurun_out/bench.log

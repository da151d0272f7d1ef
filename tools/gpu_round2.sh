#!/bin/bash
mkdir -p gpurun_out
echo "== bench"; timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench.log 2>&1; tail -3 gpurun_out/bench.log
echo "== stage times"; timeout 300 python tools/profile_stages.py c4 8192 2 2>&1 | tail -8
echo "== ncu full exchange"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:exchange -c 1 -f -o gpurun_out/prof_exchange python tools/profile_stages.py c4 2368 1 > gpurun_out/ncu_ex.log 2>&1; tail -2 gpurun_out/ncu_ex.log
echo "== ncu full taylor"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:taylor -c 1 -f -o gpurun_out/prof_taylor python tools/profile_stages.py c4 2368 1 > gpurun_out/ncu_ta.log 2>&1; tail -2 gpurun_out/ncu_ta.log
echo "== ncu full greens+gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:"greens|gemm_frag" -c 6 -f -o gpurun_out/prof_misc python tools/profile_stages.py c4 2368 1 > gpurun_out/ncu_misc.log 2>&1; tail -2 gpurun_out/ncu_misc.log
ls -la gpurun_out

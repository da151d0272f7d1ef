#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"taylor2_kernel" -c 1 -f -o gpurun_out/prof_t2 python tools/profile_stages.py c4 2368 1 > gpurun_out/ncu_t2.log 2>&1; tail -2 gpurun_out/ncu_t2.log

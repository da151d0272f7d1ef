#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"qr_kernel|exx_eri_kernel" -c 2 -f -o gpurun_out/prof_misc python tools/profile_stages.py c4 8192 1 > gpurun_out/ncu_misc.log 2>&1; tail -3 gpurun_out/ncu_misc.log

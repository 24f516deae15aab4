#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"^(eq_kernel|comp_kernel|final_kernel)$" -s 9 -c 3 -o gpurun_out/fx2_r29 -f \
    python tools/fx_bench.py 256 262144 1 > gpurun_out/r29_ncu.log 2>&1
tail -3 gpurun_out/r29_ncu.log

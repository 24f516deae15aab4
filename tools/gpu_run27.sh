#!/bin/bash
# round-1 session 2, call 1: CUDA-core pipe micro-benchmark + ncu --set full of the three FX kernels (config 3)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv | tee gpurun_out/r27_smi.log
timeout 120 tools/ubench/fp_pipes 2>&1 | tee gpurun_out/r27_ubench.log
timeout 300 python tools/fx_bench.py 256 262144 10 2>&1 | tee gpurun_out/r27_fx_bench.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:fx_ -s 9 -c 3 -o gpurun_out/fx_r27 -f \
    python tools/fx_bench.py 256 262144 1 > gpurun_out/r27_ncu.log 2>&1
tail -3 gpurun_out/r27_ncu.log

#!/bin/bash
# FX v2.2: barrier EQ with fp32 packed scan + scalar-broadcast coefficients; comp FMNMX/funnel-shift steps
mkdir -p gpurun_out
{
timeout -s KILL 300 python -m pytest tests/test_gpu_fx.py -m gpu -x -q 2>&1 | tail -5
timeout -s KILL 200 python tools/fx_bench.py 256 262144 20 2>&1 | tail -6
} | tee gpurun_out/r32.log
timeout -s KILL 600 ncu --set full --clock-control none --import-source on -k regex:"^(eq_kernel|comp_kernel|final_kernel)$" -s 9 -c 3 -o gpurun_out/fx2_r32 -f \
    python tools/fx_bench.py 256 262144 1 > gpurun_out/r32_ncu.log 2>&1
tail -2 gpurun_out/r32_ncu.log

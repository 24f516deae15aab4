#!/bin/bash
# FX v2 kernels: parity tests + A/B timing against v1
mkdir -p gpurun_out
{
timeout 600 python -m pytest tests/test_gpu_fx.py -m gpu -x -q 2>&1 | tail -15
echo "=== v2 ==="; timeout 300 python tools/fx_bench.py 256 262144 10 2>&1 | tail -6
echo "=== v1 ==="; MST_FX_IMPL=v1 timeout 300 python tools/fx_bench.py 256 262144 10 2>&1 | tail -6
} | tee gpurun_out/r28.log
timeout 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:fx2 -s 9 -c 3 --csv --log-file gpurun_out/r28_fx2_launches.csv python tools/fx_bench.py 256 262144 1 > /dev/null 2>&1
cat gpurun_out/r28_fx2_launches.csv | tail -20

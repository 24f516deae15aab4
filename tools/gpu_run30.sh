#!/bin/bash
# FX v3 (flag-chained EQ warps, fp32 packed scan; comp cp.async prefetch + balanced prefix): parity + timing + launch list
mkdir -p gpurun_out
{
timeout -s KILL 300 python -m pytest tests/test_gpu_fx.py -m gpu -x -q 2>&1 | tail -15
echo "=== v3 ==="; timeout -s KILL 200 python tools/fx_bench.py 256 262144 20 2>&1 | tail -6
} | tee gpurun_out/r30.log
timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"^(eq_kernel|comp_kernel|final_kernel)$" -s 9 -c 3 --csv --log-file gpurun_out/r30_fx2_launches.csv python tools/fx_bench.py 256 262144 1 > /dev/null 2>&1
tail -4 gpurun_out/r30_fx2_launches.csv

#!/bin/bash
# FX EQ sync-mode A/B: barrier (default) vs chain (MST_FX_EQ_SYNC=chain), parity for both
mkdir -p gpurun_out
{
echo "=== barrier ==="
timeout -s KILL 300 python -m pytest tests/test_gpu_fx.py -m gpu -x -q 2>&1 | tail -5
timeout -s KILL 200 python tools/fx_bench.py 256 262144 20 2>&1 | tail -6
echo "=== chain ==="
MST_FX_EQ_SYNC=chain timeout -s KILL 300 python -m pytest tests/test_gpu_fx.py -m gpu -x -q 2>&1 | tail -5
MST_FX_EQ_SYNC=chain timeout -s KILL 200 python tools/fx_bench.py 256 262144 20 2>&1 | tail -6
} | tee gpurun_out/r31.log
for m in barrier chain; do
MST_FX_EQ_SYNC=$m timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"^(eq_kernel|comp_kernel|final_kernel)$" -s 9 -c 3 --csv --log-file gpurun_out/r31_${m}.csv python tools/fx_bench.py 256 262144 1 > /dev/null 2>&1
done
grep -h "eq_kernel" gpurun_out/r31_barrier.csv gpurun_out/r31_chain.csv | cut -d, -f5,13,15

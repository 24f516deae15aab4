#!/bin/bash
mkdir -p gpurun_out
{
for v in "MST_TCN_PRECISION=f16f8" "MST_TCN_PRECISION=bf16x3" "MST_TCN_PRECISION=f16f8"; do env $v timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2 | head -1; done
} | tee gpurun_out/r48.log
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second \
  --clock-control none -k regex:tcn_block_umma -s 13 -c 8 --csv --log-file gpurun_out/r48_ncu.csv python tools/tcn_time.py > /dev/null 2>&1
grep -E "tcn_block" gpurun_out/r48_ncu.csv | awk -F'","' '{printf "%s ", $15}' | sed 's/"//g'; echo

#!/bin/bash
mkdir -p gpurun_out
{
timeout -s KILL 600 python -m pytest tests/test_gpu_tcn.py -m gpu -x -q 2>&1 | tail -3
for v in "MST_TCN_PAIRED=0" "MST_TCN_PAIRED=1" "MST_TCN_PAIRED=0" "MST_TCN_PAIRED=1"; do env $v timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2; done
} | tee gpurun_out/r43.log
for v in 0 1; do
MST_TCN_PAIRED=$v timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second \
  --clock-control none -k regex:tcn_block_umma -s 13 -c 13 --csv --log-file gpurun_out/r43_p$v.csv python tools/tcn_time.py > /dev/null 2>&1
grep -E "tcn_block" gpurun_out/r43_p$v.csv | awk -F'","' '{printf "%s ", $15}' | sed 's/"//g'; echo
done

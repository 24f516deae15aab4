#!/bin/bash
mkdir -p gpurun_out
{
timeout -s KILL 600 python -m pytest tests/test_gpu_tcn.py -m gpu -x -q 2>&1 | tail -3
for la in 1 0 1 0; do MST_TCN_LOOKAHEAD=$la timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2; done
} | tee gpurun_out/r42.log
for la in 1 0; do
MST_TCN_LOOKAHEAD=$la timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second \
  --clock-control none -k regex:tcn_block_umma -s 13 -c 6 --csv --log-file gpurun_out/r42_la$la.csv python tools/tcn_time.py > /dev/null 2>&1
grep -E "tcn_block" gpurun_out/r42_la$la.csv | awk -F'","' '{print $13, $15}' | tr '\n' ' '; echo
done

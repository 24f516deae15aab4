#!/bin/bash
mkdir -p gpurun_out
{
timeout -s KILL 600 python -m pytest tests/test_gpu_tcn.py -m gpu -x -q 2>&1 | tail -2
for i in 1 2; do timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2; done
} | tee gpurun_out/r50.log
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,dram__bytes_read.sum,lts__t_sector_hit_rate.pct \
  --clock-control none -k regex:tcn_block_umma -s 19 -c 7 --csv --log-file gpurun_out/r50_ncu.csv python tools/tcn_time.py > /dev/null 2>&1
grep -E "tcn_block" gpurun_out/r50_ncu.csv | awk -F'","' '{printf "%s ", $15}' | sed 's/"//g'; echo

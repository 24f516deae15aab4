#!/bin/bash
mkdir -p gpurun_out
export MST_TCN_PRECISION=f16f8
( time timeout -s KILL 2400 python -m pytest tests -m gpu -q -x --deselect tests/test_gpu_tcn_modes.py 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r45_tests_f16f8.log
timeout -s KILL 600 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 > gpurun_out/r45_bench_f16f8.json; cut -c1-330 gpurun_out/r45_bench_f16f8.json
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,l1tex__m_xbar2l1tex_read_bytes.sum \
  --clock-control none -k regex:tcn_block_umma -s 13 -c 13 --csv --log-file gpurun_out/r45_f16f8_ncu.csv python tools/tcn_time.py > /dev/null 2>&1
grep -E "tcn_block" gpurun_out/r45_f16f8_ncu.csv | awk -F'","' '{printf "%s ", $15}' | sed 's/"//g'; echo

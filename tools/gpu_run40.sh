#!/bin/bash
mkdir -p gpurun_out
{
for dbg in 4 0; do MST_TCN_PRECISION=f16f8 MST_TCN_DBG=$dbg timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2; done
for dbg in 4; do MST_TCN_PIPE=2 MST_TCN_DBG=$dbg timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2; done
timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2
} | tee gpurun_out/r40.log

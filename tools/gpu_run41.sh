#!/bin/bash
mkdir -p gpurun_out
{
for dbg in 5 6 14 22 12 20; do MST_TCN_PRECISION=f16f8 MST_TCN_DBG=$dbg timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2 | head -1; done
} | tee gpurun_out/r41.log

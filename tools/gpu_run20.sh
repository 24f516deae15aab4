#!/bin/bash
mkdir -p gpurun_out
export MST_TCN_PRECISION=f16f8 MST_TCN_MULTICAST=0
for dbg in 14 22 7; do MST_TCN_DBG=$dbg timeout 200 python tools/tcn_time.py 2>&1 | tail -1; done | tee gpurun_out/dbg20.log

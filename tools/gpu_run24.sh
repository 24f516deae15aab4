#!/bin/bash
mkdir -p gpurun_out
{
for d in 6 70 102 7 71; do MST_TCN_PIPE=2 MST_TCN_DBG=$d timeout 200 python tools/tcn_time.py 2>&1 | tail -1; done
} | tee gpurun_out/dbg24.log

#!/bin/bash
mkdir -p gpurun_out
{
for d in 0 6 38 32 2 36; do MST_TCN_PIPE=2 MST_TCN_DBG=$d timeout 200 python tools/tcn_time.py 2>&1 | tail -1; done
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit --format=csv
} | tee gpurun_out/dbg23.log

#!/bin/bash
mkdir -p gpurun_out
for o in 0 1 0 1; do MST_TCN_MMA_ORDER=$o timeout 200 python tools/tcn_time.py 2>&1 | tail -1 | sed "s/^/order=$o /"; done | tee gpurun_out/dbg21.log

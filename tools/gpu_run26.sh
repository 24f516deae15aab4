#!/bin/bash
mkdir -p gpurun_out
{
for f in 0 1 0 1; do MST_TCN_NOFENCE=$f timeout 200 python tools/tcn_time.py 2>&1 | tail -1 | sed "s/^/pipe1 nofence=$f /"; done
for d in 0 256 6 262 358; do MST_TCN_PIPE=2 MST_TCN_DBG=$d timeout 200 python tools/tcn_time.py 2>&1 | tail -1; done
MST_TCN_NOFENCE=1 timeout 600 python -m pytest tests/test_gpu_tcn.py -m gpu -x -q 2>&1 | tail -3
} | tee gpurun_out/dbg26.log

#!/bin/bash
mkdir -p gpurun_out
{
MST_TCN_PIPE=2 timeout 600 python -m pytest tests/test_gpu_tcn.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -5
for p in 1 2 1 2; do MST_TCN_PIPE=$p timeout 200 python tools/tcn_time.py 2>&1 | tail -1; done
MST_TCN_PIPE=2 MST_TCN_MULTICAST=1 timeout 200 python tools/tcn_time.py 2>&1 | tail -1
MST_TCN_PIPE=2 MST_TCN_DBG=1 timeout 200 python tools/tcn_time.py 2>&1 | tail -1
MST_TCN_PIPE=2 MST_TCN_DBG=4 timeout 200 python tools/tcn_time.py 2>&1 | tail -1
} | tee gpurun_out/dbg22.log

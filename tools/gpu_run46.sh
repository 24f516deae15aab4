#!/bin/bash
mkdir -p gpurun_out
{
timeout -s KILL 600 python -m pytest tests/test_gpu_tcn.py -m gpu -x -q 2>&1 | tail -3
MST_TCN_LOOKAHEAD=1 timeout -s KILL 600 python -m pytest tests/test_gpu_tcn.py -m gpu -x -q 2>&1 | tail -3
for v in "MST_TCN_LOOKAHEAD=0" "MST_TCN_LOOKAHEAD=1" "MST_TCN_LOOKAHEAD=0" "MST_TCN_LOOKAHEAD=1"; do env $v timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2 | head -1; done
} | tee gpurun_out/r46.log

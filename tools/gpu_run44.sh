#!/bin/bash
mkdir -p gpurun_out
{
echo "== parity f16f8 (single ring, paired)"; MST_TCN_PRECISION=f16f8 timeout -s KILL 600 python -m pytest tests/test_gpu_tcn.py -m gpu -x -q 2>&1 | tail -4
echo "== parity f16f8 unpaired"; MST_TCN_PRECISION=f16f8 MST_TCN_PAIRED=0 timeout -s KILL 600 python -m pytest tests/test_gpu_tcn.py -m gpu -x -q 2>&1 | tail -3
for v in "MST_TCN_PRECISION=f16f8" "MST_TCN_PAIRED=1" "MST_TCN_PRECISION=f16f8 MST_TCN_PAIRED=0" "MST_TCN_PRECISION=f16f8"; do env $v timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2; done
} | tee gpurun_out/r44.log

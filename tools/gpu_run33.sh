#!/bin/bash
# TCN paired sub-tiles (d >= 128): parity + A/B timing; FX with the EQ L2 prefetch
mkdir -p gpurun_out
{
timeout -s KILL 900 python -m pytest tests/test_gpu_tcn.py -m gpu -x -q 2>&1 | tail -8
for p in 1 0 1 0; do MST_TCN_PAIRED=$p timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2; done
timeout -s KILL 300 python -m pytest tests/test_gpu_fx.py -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 200 python tools/fx_bench.py 256 262144 20 2>&1 | head -1
} | tee gpurun_out/r33.log

#!/bin/bash
mkdir -p gpurun_out
export MST_TCN_PRECISION=f16f8 MST_TCN_MULTICAST=0
for dbg in 0 6 3; do MST_TCN_DBG=$dbg timeout 200 python tools/tcn_time.py 2>&1 | tail -1; done | tee gpurun_out/dbg19.log
MST_TCN_MULTICAST=1 MST_TCN_DBG=0 timeout 200 python tools/tcn_time.py 2>&1 | tail -1 | tee -a gpurun_out/dbg19.log
timeout 600 python -m pytest tests/test_gpu_tcn.py -q --tb=line -x 2>&1 | tail -3 | tee -a gpurun_out/dbg19.log

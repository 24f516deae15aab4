#!/bin/bash
mkdir -p gpurun_out
{ timeout -s KILL 200 python tools/tcn_error_report.py 2>&1 | tail -3; MST_TCN_PRECISION=bf16x3 timeout -s KILL 200 python tools/tcn_error_report.py 2>&1 | tail -3; } | tee gpurun_out/r52_errors.log

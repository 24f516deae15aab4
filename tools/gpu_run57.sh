#!/bin/bash
mkdir -p gpurun_out
timeout -s KILL 300 python -m pytest tests/test_gpu_e2e.py -m gpu -q -k "feature_extraction or smoke or device_io" 2>&1 | tail -3 | tee gpurun_out/r57.log

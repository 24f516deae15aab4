#!/bin/bash
# full GPU suite + headline bench (state: FX v2.2, TCN paired opt-in)
mkdir -p gpurun_out
( time timeout -s KILL 2400 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) 2>&1 | tee gpurun_out/r35_tests.log
timeout -s KILL 900 python bench.py 2>gpurun_out/r35_bench.err | tee gpurun_out/r35_bench.json | cut -c1-600
tail -3 gpurun_out/r35_bench.err

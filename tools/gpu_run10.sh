#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log; }
run fx10     600 python -m pytest tests/test_gpu_fx.py -q --tb=short

run fxbench10 600 python tools/fx_bench.py 256 262144 10


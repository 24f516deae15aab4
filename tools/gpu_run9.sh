#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 8 gpurun_out/$name.log; }
run fx9      600 python -m pytest tests/test_gpu_fx.py -q --tb=short
run fxbench9 600 python tools/fx_bench.py 256 262144 10
run fxlaunch9 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/fx_launches_r01b.csv python tools/fx_bench.py 256 262144 1

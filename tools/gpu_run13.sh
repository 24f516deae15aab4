#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 5 gpurun_out/$name.log; }
run all13    1500 python -m pytest tests -q -m gpu --tb=short
run bench13  900 python bench.py
run launches13 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 220 -c 60 --csv --log-file gpurun_out/launches_r01d.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline

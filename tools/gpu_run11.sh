#!/bin/bash
# full validation: every GPU test + smoke + bench + launch lists (profiles)
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 4 gpurun_out/$name.log; }
run all11    1500 python -m pytest tests -q -m gpu --tb=short
run bench11  900 python bench.py
run launches11 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 220 -c 60 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
run fxlaunch11 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 6 -c 12 --csv --log-file gpurun_out/fx_launches_r01c.csv python tools/fx_bench.py 256 262144 1

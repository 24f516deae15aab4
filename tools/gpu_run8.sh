#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 8 gpurun_out/$name.log; }
run enc8     900 python -m pytest tests/test_gpu_encoder.py -q --tb=short
run e2e8     900 python -m pytest tests/test_gpu_e2e.py -q --tb=short -s -k "stem_style or smoke"
run bench8   600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run launches8 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 200 -c 120 --csv --log-file gpurun_out/launches_r01b.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline

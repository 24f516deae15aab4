#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log; }
run enc12    900 python -m pytest tests/test_gpu_encoder.py -q --tb=short -s -k "full_encoder or golden"
run e2e12    900 python -m pytest tests/test_gpu_e2e.py -q --tb=short -s
run bench12  600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline

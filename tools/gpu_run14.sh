#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log; }
export MST_TCN_PRECISION=f16f8
run dbg14    300 python tools/tcn_debug.py 1 512 1
run tcn14    900 python -m pytest tests/test_gpu_tcn.py -q --tb=line
run bench14  600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline

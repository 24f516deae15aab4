#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log; }
export MST_TCN_PRECISION=f16f8
run tcn15    900 python -m pytest tests/test_gpu_tcn.py -q --tb=line
run bench15  600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run ncu15    900 ncu --set full --clock-control none --import-source on -k regex:block_kernel -s 5 -c 1 -o gpurun_out/f8_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline

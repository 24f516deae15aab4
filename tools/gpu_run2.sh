#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 12 gpurun_out/$name.log; }
run e2e      900 python -m pytest tests/test_gpu_e2e.py -q --tb=short
run fxbench  600 python tools/fx_bench.py 256 262144 5
# launch list of one bench step (cold-cache, serialised: shares only)
run launches 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_r01.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline
run fxlaunch 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/fx_launches_r01.csv python tools/fx_bench.py 256 262144 1
# one full-set capture of the dominant kernel (3 launches after warm-up)
run ncufull 1200 ncu --set full --clock-control none --import-source on -k regex:tcn_block_umma -s 13 -c 3 -o gpurun_out/umma_r01 python bench.py --steps 1 --warmup 3 --no-cpu-baseline

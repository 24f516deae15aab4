#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log; }
export MST_TCN_KCHUNK=32
run tcn32    900 python -m pytest tests/test_gpu_tcn.py -q --tb=short -x
run bench32  600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
export MST_TCN_KCHUNK=64
run bench64  600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run refarm   900 python bench.py --impl reference --steps 2 --warmup 1

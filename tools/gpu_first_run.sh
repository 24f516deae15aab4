#!/bin/bash
# First GPU bring-up: every stage in its own process with a timeout so a hung kernel cannot eat the box.
mkdir -p gpurun_out
nvidia-smi > gpurun_out/nvidia_smi.txt 2>&1
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 25 gpurun_out/$name.log; }
run enc      600 python -m pytest tests/test_gpu_encoder.py -q -x --tb=short
run tcndbg1  300 python tools/tcn_debug.py 1 512 1
run tcndbg0  300 python tools/tcn_debug.py 0 512 1
run tcn      900 python -m pytest tests/test_gpu_tcn.py -q --tb=short
run fx       600 python -m pytest tests/test_gpu_fx.py -q --tb=short
run e2e      900 python -m pytest tests/test_gpu_e2e.py -q --tb=short
run bench    900 python bench.py --steps 3 --warmup 3

#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/n2_gpus.txt 2>&1
echo "=== bench N=2"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n2.log 2>&1; echo "exit $?" >> gpurun_out/bench_n2.log; tail -n 5 gpurun_out/bench_n2.log
echo "=== reference arm N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/ref_n2.log 2>&1; echo "exit $?" >> gpurun_out/ref_n2.log; tail -n 3 gpurun_out/ref_n2.log

#!/bin/bash
# 2 GPUs: NCCL shard layer (both modes) bit-identical to one process; bench at N=2 and N=1 on the same box
mkdir -p gpurun_out
{
timeout -s KILL 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/nccl_interp_check.py 2>&1 | grep -E "rank|Error|error" | head
timeout -s KILL 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-420
timeout -s KILL 900 python bench.py --gpus 1 --steps 10 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-420
} | tee gpurun_out/r38_n2.log

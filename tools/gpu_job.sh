#!/bin/bash
# One parameterised GPU job (replaces the numbered one-shot scripts of round 1).  usage: tools/gpu_job.sh <what> ...
#   tests            pytest -m gpu
#   ablate           default build vs the MST_TCN_ABLATE side builds (build.py --variant), tools/tcn_time.py each
#   tcntests         the TCN parity files only
#   ab               tools/tcn_time.py for build/base (previous commit), the product library and $MST_AB_VARIANTS side builds
#   nccl / benchn    tools/nccl_check.py / bench.py under torchrun with $MST_NPROC ranks (gpurun --gpus N)
#   bench            python bench.py
# Everything is written under gpurun_out/.
set -u
mkdir -p gpurun_out
for what in "$@"; do
case "$what" in
  tests)  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/tests.log; tail -5 gpurun_out/tests.log ;;
  ablate) { timeout 300 python tools/tcn_time.py f16f8
            for m in 1 2 3 4 6 7; do MST_DEV_LIB=music_mixing_style_transfer_b200/build/abl$m/libmst_b200.so timeout 300 python tools/tcn_time.py f16f8; done
            timeout 300 python tools/tcn_time.py bf16x3; } 2>&1 | grep -v Warning > gpurun_out/ablate.log; cat gpurun_out/ablate.log ;;
  tcntests) timeout 900 python -m pytest tests/test_gpu_tcn.py tests/test_gpu_tcn_modes.py tests/test_gpu_fullsize.py -m gpu -x -q 2>&1 | tail -15 > gpurun_out/tcntests.log; tail -4 gpurun_out/tcntests.log ;;
  ab)     { for v in base "" $MST_AB_VARIANTS; do
              lib=""; [ -n "$v" ] && lib=music_mixing_style_transfer_b200/build/$v/libmst_b200.so
              MST_DEV_LIB=$lib timeout 300 python tools/tcn_time.py f16f8 32 ${MST_AB_REPS:-4}
            done; } 2>&1 | grep -v Warning > gpurun_out/ab.log; cat gpurun_out/ab.log ;;
  nccl)   timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${MST_NPROC:-2} --master-addr 127.0.0.1 --master-port 29511 tools/nccl_check.py 2>&1 | grep -E "^rank|Error|error" | tail -20 > gpurun_out/nccl.log; cat gpurun_out/nccl.log ;;
  benchn) timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node ${MST_NPROC:-2} --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus ${MST_NPROC:-2} --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench_n${MST_NPROC:-2}.json 2> gpurun_out/bench_n.err; tail -c 2500 gpurun_out/bench_n${MST_NPROC:-2}.json; tail -3 gpurun_out/bench_n.err ;;
  bench)  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json ;;
esac
done

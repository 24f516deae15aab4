#!/bin/bash
# One parameterised GPU job (replaces the numbered one-shot scripts of round 1).  usage: tools/gpu_job.sh <what> ...
#   tests            pytest -m gpu
#   ablate           default build vs the MST_TCN_ABLATE side builds (build.py --variant), tools/tcn_time.py each
#   bench [args]     python bench.py args
# Everything is written under gpurun_out/.
set -u
mkdir -p gpurun_out
for what in "$@"; do
case "$what" in
  tests)  timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -25 > gpurun_out/tests.log; tail -5 gpurun_out/tests.log ;;
  ablate) { timeout 300 python tools/tcn_time.py f16f8
            for m in 1 2 3 4 6 7; do MST_DEV_LIB=music_mixing_style_transfer_b200/build/abl$m/libmst_b200.so timeout 300 python tools/tcn_time.py f16f8; done
            timeout 300 python tools/tcn_time.py bf16x3; } 2>&1 | grep -v Warning > gpurun_out/ablate.log; cat gpurun_out/ablate.log ;;
  bench)  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json ;;
esac
done

#!/bin/bash
# One parameterised GPU job (replaces the numbered one-shot scripts of round 1).  usage: tools/gpu_job.sh <what> ...
#   tests            pytest -m gpu
#   ablate           default build vs the MST_TCN_ABLATE side builds (build.py --variant), tools/tcn_time.py each
#   tcntests         the TCN parity files only
#   ab               tools/tcn_time.py for build/base (previous commit), the product library and $MST_AB_VARIANTS side builds
#   nccl / benchn    tools/nccl_check.py / bench.py under torchrun with $MST_NPROC ranks (gpurun --gpus N)
#   bench            python bench.py
#   profile          the round's ncu evidence (launch lists, TCN metrics + one --set full capture, FX and normaliser kernels)
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
  profile) # ncu evidence of a round: launch list of bench steps, per-launch TCN metrics, one --set full capture, FX / new-kernel lists
          TAG=${MST_PROFILE_TAG:-r02c}
          timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/${TAG}_bench_launches.csv \
              python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-extra > gpurun_out/${TAG}_bench_under_ncu.log 2>&1
          timeout 600 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sector_hit_rate.pct \
              --clock-control none -k regex:tcn_block_umma -s 13 -c 13 --csv --log-file gpurun_out/${TAG}_tcn_ncu.csv python tools/tcn_time.py f16f8 32 2 > /dev/null 2>&1
          timeout 600 ncu --set full --clock-control none --import-source on -k regex:tcn_block_umma -s 21 -c 1 -f -o gpurun_out/${TAG}_tcn_full python tools/tcn_time.py f16f8 32 2 > /dev/null 2>&1
          ncu -i gpurun_out/${TAG}_tcn_full.ncu-rep --page raw --csv > gpurun_out/${TAG}_tcn_ncu_full.csv 2>/dev/null
          timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,sm__warps_active.avg.pct_of_peak_sustained_active \
              --clock-control none -k regex:"eq_kernel|comp_kernel|final_kernel" -s 9 -c 3 --csv --log-file gpurun_out/${TAG}_fx_launches.csv python tools/fx_bench.py 256 262144 1 > /dev/null 2>&1
          timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none \
              -k regex:"fft_|mag_acc|fir64|conv_|network_kernel|mix_kernel|row_absmax" --csv --log-file gpurun_out/${TAG}_new_kernels.csv python tools/norm_bench.py > gpurun_out/${TAG}_norm_bench_under_ncu.log 2>&1
          timeout 300 python tools/norm_bench.py > gpurun_out/${TAG}_norm_bench.log 2>&1; cat gpurun_out/${TAG}_norm_bench.log
          rm -f gpurun_out/${TAG}_tcn_full.ncu-rep
          ls -la gpurun_out | tail -12 ;;
  bench)  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json ;;
esac
done

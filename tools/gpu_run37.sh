#!/bin/bash
# end-of-session evidence: bench line, ncu launch lists (bench step, FX chain), FX per-kernel --set full summary metrics
mkdir -p gpurun_out
timeout -s KILL 900 python bench.py 2>gpurun_out/r37_bench.err > gpurun_out/r37_bench.json; cut -c1-300 gpurun_out/r37_bench.json
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 220 -c 60 --csv --log-file gpurun_out/r37_bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed \
    --clock-control none -k regex:"^(eq_kernel|comp_kernel|final_kernel)$" -s 9 -c 3 --csv --log-file gpurun_out/r37_fx_launches.csv python tools/fx_bench.py 256 262144 1 > /dev/null 2>&1
timeout -s KILL 300 python tools/fx_bench.py 256 262144 20 2>&1 | tee gpurun_out/r37_fx_bench.log | head -3
wc -l gpurun_out/r37_bench_launches.csv gpurun_out/r37_fx_launches.csv

#!/bin/bash
# final state of session 2: full GPU suite, bench line, launch list, TCN ncu metrics (default = f16f8, paired, convergent issuer)
mkdir -p gpurun_out
( time timeout -s KILL 2400 python -m pytest tests -m gpu -q 2>&1 | tail -8 ) 2>&1 | tee gpurun_out/r49_tests.log
timeout -s KILL 900 python bench.py 2>gpurun_out/r49_bench.err > gpurun_out/r49_bench.json; cut -c1-260 gpurun_out/r49_bench.json
timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2 | tee gpurun_out/r49_tcn_time.log
timeout -s KILL 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 225 -c 45 --csv --log-file gpurun_out/r49_bench_launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sector_hit_rate.pct \
  --clock-control none -k regex:tcn_block_umma -s 13 -c 13 --csv --log-file gpurun_out/r49_tcn_ncu.csv python tools/tcn_time.py > /dev/null 2>&1
wc -l gpurun_out/r49_bench_launches.csv gpurun_out/r49_tcn_ncu.csv

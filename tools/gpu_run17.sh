#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; echo "=== $name"; timeout "$@" > gpurun_out/$name.log 2>&1; echo "exit $?" >> gpurun_out/$name.log; tail -n 6 gpurun_out/$name.log; }
export MST_TCN_PRECISION=f16f8
run dbg17    200 python tools/tcn_debug.py 1 4099 2
run tcn17    600 python -m pytest tests/test_gpu_tcn.py -q --tb=line -x
run bench17  600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline
run ncu17    900 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__cycles_elapsed.avg.per_second,l1tex__m_xbar2l1tex_read_bytes.sum --clock-control none -k regex:block_kernel -s 5 -c 1 --csv --log-file gpurun_out/f8_metrics.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline

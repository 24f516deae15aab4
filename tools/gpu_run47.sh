#!/bin/bash
mkdir -p gpurun_out
{
for v in "MST_TCN_PRECISION=f16f8" "MST_TCN_PRECISION=bf16x3" "MST_TCN_PRECISION=f16f8" "MST_TCN_PRECISION=bf16x3"; do env $v timeout -s KILL 200 python tools/tcn_time.py 2>&1 | tail -2 | head -1; done
} | tee gpurun_out/r47.log
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__m_xbar2l1tex_read_bytes.sum \
  --clock-control none -k regex:tcn_block_umma -s 13 -c 13 --csv --log-file gpurun_out/r47_f16f8_ncu.csv python tools/tcn_time.py > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r47_f16f8_ncu.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
d={}
for r in rows[h+1:]: d.setdefault(r[0],{})[r[12].split('.')[0][-22:]]=r[14]
for k,v in d.items(): print(k, v)
PY

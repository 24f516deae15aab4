#!/bin/bash
mkdir -p gpurun_out
{
timeout -s KILL 900 python -m pytest tests/test_gpu_encoder.py tests/test_gpu_e2e.py -m gpu -x -q 2>&1 | tail -3
timeout -s KILL 600 python bench.py --no-cpu-baseline 2>/dev/null | tail -1 | cut -c1-240
} | tee gpurun_out/r54.log
timeout -s KILL 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:enc_conv_umma -s 18 -c 18 --csv --log-file gpurun_out/r54_enc.csv python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r54_enc.csv')))
h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
t=[float(r[-1])/1e6 for r in rows[h+1:]]
print('enc_conv_umma launches', len(t), 'total ms %.3f'%sum(t))
PY

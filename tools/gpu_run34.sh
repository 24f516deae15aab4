#!/bin/bash
# TCN paired vs unpaired under ncu (isolated launches, no sustained power cap): duration, tensor pipe, L2->SM bytes, clock
mkdir -p gpurun_out
for p in 1 0; do
MST_TCN_PAIRED=$p timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.avg.per_second,lts__t_sector_hit_rate.pct,dram__bytes_read.sum \
  --clock-control none -k regex:tcn_block_umma -s 13 -c 13 --csv --log-file gpurun_out/r34_paired$p.csv python tools/tcn_time.py > /dev/null 2>&1
done
python - <<'PY'
import csv
for p in (1,0):
    rows=list(csv.reader(open(f'gpurun_out/r34_paired{p}.csv')))
    h=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
    d={}
    for r in rows[h+1:]:
        d.setdefault(r[0],{})[r[12]]=r[14]
    print('paired',p)
    for k,v in d.items():
        print(k, ' '.join(f"{n.split('.')[0][-22:]}={x}" for n,x in v.items()))
PY

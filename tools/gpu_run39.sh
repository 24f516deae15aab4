#!/bin/bash
# f16f8 (2 tensor units) and pipe-2 variants under ncu: duration, tensor activity, SM clock, operand bytes -- are they power-bound too?
mkdir -p gpurun_out
for v in "MST_TCN_PRECISION=f16f8" "MST_TCN_PIPE=2"; do
env $v timeout -s KILL 400 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,l1tex__m_xbar2l1tex_read_bytes.sum,sm__cycles_elapsed.avg.per_second,sm__inst_executed_pipe_tensor.sum,smsp__inst_executed.sum \
  --clock-control none -k regex:"block_kernel|tcn_block_umma" -s 13 -c 13 --csv --log-file gpurun_out/r39_${v#*=}.csv python tools/tcn_time.py > gpurun_out/r39_${v#*=}.log 2>&1
done
python - <<'PY'
import csv,glob
for f in sorted(glob.glob('gpurun_out/r39_*.csv')):
    rows=list(csv.reader(open(f)))
    hs=[i for i,r in enumerate(rows) if r and r[0]=='ID']
    if not hs: print(f,'no data'); continue
    d={}
    for r in rows[hs[0]+1:]:
        d.setdefault(r[0],{'k':r[4][:40]})[r[12]]=r[14]
    print(f)
    for k,v in list(d.items())[:13]:
        print(k, v['k'], ' '.join(f"{n.split('.')[0][-20:]}={x}" for n,x in v.items() if n!='k'))
PY

#!/bin/bash
# per-launch DRAM bytes / time / clock / tensor activity of the 13 dilated launches for the product library and side builds
mkdir -p gpurun_out
for v in "" $MST_AB_VARIANTS; do
  lib=""; [ -n "$v" ] && lib=music_mixing_style_transfer_b200/build/$v/libmst_b200.so
  MST_DEV_LIB=$lib timeout 300 ncu --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second,dram__bytes_read.sum,l1tex__m_xbar2l1tex_read_bytes.sum,lts__t_sector_hit_rate.pct \
      --clock-control none -k regex:tcn_block_umma -s 13 -c 13 --csv --log-file gpurun_out/ncu_var_${v:-product}.csv python tools/tcn_time.py f16f8 32 2 > /dev/null 2>&1
  python - "$v" <<'PY'
import csv, io, sys, collections
v = sys.argv[1] or "product"
txt = open(f"gpurun_out/ncu_var_{v}.csv").read()
rows = list(csv.DictReader(io.StringIO(txt[txt.find('"ID"'):])))
by = collections.OrderedDict()
for r in rows:
    by.setdefault(r['ID'], {})[r['Metric Name']] = float(r['Metric Value'].replace(',', ''))
print(v, "| ms:", " ".join("%.2f" % (m['gpu__time_duration.sum'] / 1e6) for m in by.values()))
print(v, "| dramR GB:", " ".join("%.1f" % (m['dram__bytes_read.sum'] / 1e9) for m in by.values()))
print(v, "| GHz:", " ".join("%.2f" % (m['sm__cycles_elapsed.avg.per_second'] / 1e9) for m in by.values()))
print(v, "| tensor%:", " ".join("%.0f" % m['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'] for m in by.values()))
PY
done

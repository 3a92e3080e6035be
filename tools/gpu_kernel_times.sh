#!/bin/bash
# usage: tools/gpu_kernel_times.sh "<bench args>"  -- per-kernel gpu__time_duration of one timed step (ncu launch list)
a="$1"
python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 1 $a > gpurun_out/kt_plain.log 2>&1 || { echo FAIL; tail -3 gpurun_out/kt_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/kt.csv python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 1 $a > gpurun_out/kt_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/kt.csv')) if len(r)>12 and r[0]!='ID']
from collections import OrderedDict
d=OrderedDict()
for r in rows:
    d.setdefault((int(r[0]), r[4].split('(')[0][:40], r[8]), {})[r[12]]=r[14]
seen=set()
for (i,k,g),v in list(d.items())[::-1]:
    if 'btl::' not in k or 'synth' in k or k in seen: continue
    seen.add(k)
    print("%-42s grid %-12s %7.3f ms  rd %6.2f GB wr %6.2f GB  inst %6.0f M issue %4.1f%% warps %4.1f%%" % (k, g, float(v['gpu__time_duration.sum'])/1e6, float(v['dram__bytes_read.sum'])/1e9, float(v['dram__bytes_write.sum'])/1e9, float(v['smsp__inst_executed.sum'])/1e6, float(v['smsp__issue_active.avg.pct_of_peak_sustained_active']), float(v['sm__warps_active.avg.pct_of_peak_sustained_active'])))
PY

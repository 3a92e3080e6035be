#!/bin/bash
# usage: tools/gpu_kt_generic.sh <skip> <count> <cmd...>  -- per-kernel time table for an arbitrary command
skip=$1; cnt=$2; shift 2
"$@" > gpurun_out/ktg_plain.log 2>&1 || { echo FAIL; tail -3 gpurun_out/ktg_plain.log; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -s $skip -c $cnt --csv --log-file gpurun_out/ktg.csv "$@" > gpurun_out/ktg_ncu.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/ktg.csv')) if len(r)>12 and r[0]!='ID']
from collections import OrderedDict
d=OrderedDict()
for r in rows:
    d.setdefault((int(r[0]), r[4].split('(')[0][:44], r[8]), {})[r[12]]=r[14]
agg=OrderedDict()
for (i,k,g),v in d.items():
    a=agg.setdefault(k,[0,0.0,0.0,0.0,0.0])
    a[0]+=1; a[1]+=float(v['gpu__time_duration.sum'])/1e6; a[2]+=float(v['dram__bytes_read.sum'])/1e9; a[3]+=float(v['dram__bytes_write.sum'])/1e9; a[4]+=float(v['smsp__inst_executed.sum'])/1e6
for k,a in agg.items():
    print("%-46s n=%4d  total %8.3f ms  avg %7.3f ms  rd %6.2f GB wr %6.2f GB inst %7.0f M" % (k,a[0],a[1],a[1]/a[0],a[2],a[3],a[4]))
PY

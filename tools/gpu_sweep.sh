#!/bin/bash
# usage: tools/gpu_sweep.sh "<bench args 1>" "<bench args 2>" ...   (runs on the GPU box; prints one summary line each)
for a in "$@"; do
  python bench.py --no-e2e --no-cpu-baseline --steps 6 --warmup 3 $a > gpurun_out/sweep.json 2> gpurun_out/sweep.err || { echo "FAIL $a"; tail -3 gpurun_out/sweep.err; continue; }
  python - "$a" <<'PY'
import json,sys
d=json.load(open("gpurun_out/sweep.json"))
print("%-40s value %.2f insert %.2f (%.2f ms, frac %.3f) query %.2f (%.2f ms, frac %.3f) launches %d" % (sys.argv[1], d["value"], d["insert_gkmers_s"], d["roofline_build"]["launch_ms"], d["roofline_build"]["frac"], d["query_gkmers_s"], d["roofline"]["launch_ms"], d["roofline"]["frac"], d["gpu_launches"]))
PY
done

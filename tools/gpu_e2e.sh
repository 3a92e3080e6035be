#!/bin/bash
for a in "$@"; do
  python bench.py --no-cpu-baseline --steps 12 --warmup 3 $a > gpurun_out/sweep.json 2> gpurun_out/sweep.err || { echo "FAIL $a"; tail -3 gpurun_out/sweep.err; continue; }
  python - "$a" <<'PY'
import json,sys
d=json.load(open("gpurun_out/sweep.json"))
print("%-44s value %.2f e2e %.2f e2e_sync %.2f insert %.2f query %.2f" % (sys.argv[1], d["value"], d["e2e"]["value"], d["e2e_sync"]["value"], d["insert_gkmers_s"], d["query_gkmers_s"]))
PY
done

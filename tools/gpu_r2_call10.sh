#!/bin/bash
# round 2, GPU call 10 (1 GPU): pass-2 prefetch variants (next partition / own partition / none) and 1024 partitions for the
# 16 GB filters, where the probe kernel's DRAM reads are 1.7x the items + one pass over the filter
mkdir -p gpurun_out
SKIP_BUILD=1 THRESHOLD=1 timeout 500 python tools/r2_sweep.py cfg4 "" "bin_prefetch=2" "bin_prefetch=1" "bin_max_parts=1024" "bin_max_parts=1024,bin_prefetch=1" "bin_max_parts=1024,bin_prefetch=2" \
  > gpurun_out/sweep3_cfg4.jsonl 2> gpurun_out/sweep3_cfg4.err
BUILD_REPS=16 timeout 500 python tools/r2_sweep.py cfg5b "" "bin_prefetch=2" "bin_prefetch=0" "bin_max_parts=1024" "bin_max_parts=1024,bin_prefetch=1" "bin_max_parts=1024,bin_prefetch=2" \
  > gpurun_out/sweep3_cfg5b.jsonl 2> gpurun_out/sweep3_cfg5b.err
timeout 300 python tools/r2_sweep.py cfg2 "" "bin_prefetch=2" "bin_prefetch=0" > gpurun_out/sweep3_cfg2.jsonl 2> gpurun_out/sweep3_cfg2.err
timeout 300 python tools/r2_sweep.py cfg3 "" "bin_prefetch=2" > gpurun_out/sweep3_cfg3.jsonl 2> gpurun_out/sweep3_cfg3.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/sweep3_*.jsonl')):
    for l in open(f):
        d = json.loads(l)
        print(d['config'], d['options'], 'build %.2f' % d.get('build_gkmers_s', 0), 'query %.2f' % d['query_gkmers_s'], 'q_ms %.2f' % d['query_ms'])
PY
tail -3 gpurun_out/sweep3_*.err

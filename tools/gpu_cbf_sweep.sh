#!/bin/bash
# usage: tools/gpu_cbf_sweep.sh "<opt=val ...>" ...   -- cfg4 (CountingBloomFilter) kernel throughput per option set
for a in "$@"; do
  CONFIGS=cfg4 python tools/bench_configs.py $a 2> gpurun_out/cbf_sweep.err | python -c "
import json,sys
for l in sys.stdin:
    d=json.loads(l); print('%-60s insert %.3f Gk/s (%.2f ms) query %.2f deferred/rounds %s' % (' '.join(d['options']), d['insert_gkmers_s'], d['insert_ms'], d['query_gkmers_s'], d['ordered_deferred_rounds']))
" || { echo "FAIL $a"; tail -3 gpurun_out/cbf_sweep.err; }
done

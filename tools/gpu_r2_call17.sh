#!/bin/bash
# round 2, GPU call 17 (N GPUs): merge pipelined behind pass 2 of the build (btlbf_filter_flush_parts + btlbf_merge_peers_range)
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "2" ]; then
  timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_multi17.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_multi17.log | cut -c1-300
fi
short="--gpus $N --steps 8 --warmup 3 --no-configs --no-e2e --no-cpu-baseline"
for c in ${CHUNKS:-0 2 4 8}; do
  timeout 400 $TR --master-port 2981$c bench.py $short --merge-chunks $c > gpurun_out/bench_r2_n${N}_chunks$c.json 2> gpurun_out/bench_r2_n${N}_chunks$c.err; echo "chunks $c rc=$?"
done
python - <<PY
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_r2_n${N}_chunks*.json')):
    txt = [l for l in open(f).read().splitlines() if l.startswith('{')]
    if not txt:
        print(f, 'no json', open(f.replace('.json', '.err')).read()[-600:]); continue
    d = json.loads(txt[-1]); m = d['merge']; j = d['job']
    print(f, 'value %.1f build_ms %.2f merge_ms %.2f query_ms %.2f | job %.1f build %.2f merge %.2f query %.2f pop_ok %s' % (d['value'], d['build_ms'], d['merge_ms'], d['query_ms'], j['gkmers_s'], j['build_ms'], j['merge_ms'], j['query_ms'], j['identical_popcount_on_all_ranks']), j['popcount'], d['merge_check']['merge_parity'])
PY

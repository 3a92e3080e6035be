#!/bin/bash
# round 2, GPU call 8 (8 GPUs): multi-GPU worker at world 4, the driver's N=8 bench command, the merge kernels side by side
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -s > gpurun_out/pytest_multi8.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi8.log
grep -a "MULTI-GPU\|passed\|failed\|rc=" gpurun_out/pytest_multi8.log | tail -4
timeout 900 $TR --master-port 29611 bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/bench_r2_n8.json 2> gpurun_out/bench_r2_n8.err; echo "bench n8 rc=$?"
tail -c 800 gpurun_out/bench_r2_n8.err
short="--gpus 8 --steps 8 --warmup 3 --no-configs --no-e2e --no-job --no-cpu-baseline"
timeout 300 $TR --master-port 29612 bench.py $short --merge peer > gpurun_out/bench_r2_n8_peer.json 2> gpurun_out/bench_r2_n8_peer.err
timeout 300 $TR --master-port 29613 bench.py $short --merge multimem > gpurun_out/bench_r2_n8_mm.json 2> gpurun_out/bench_r2_n8_mm.err
timeout 300 $TR --master-port 29614 bench.py $short --merge multimem --opt mm_unroll=8 > gpurun_out/bench_r2_n8_mm8.json 2> gpurun_out/bench_r2_n8_mm8.err
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_r2_n8*.json')):
    txt = [l for l in open(f).read().splitlines() if l.startswith('{')]
    if not txt:
        print(f, 'no json'); continue
    d = json.loads(txt[-1]); m = d.get('merge', {})
    print(f, 'value %.1f' % d.get('value', 0), 'per_gpu', d.get('per_gpu_rate'), 'merge', m.get('kind'), m.get('ms'), m.get('calibration_ms'), 'e2e', d.get('e2e', {}).get('value'), 'pk', d.get('e2e_packed', {}).get('value'), 'job', (d.get('job') or {}).get('gkmers_s'), (d.get('job') or {}).get('merge_ms'))
    for n, c in (d.get('configs') or {}).items():
        print('   ', n, c.get('error') or ('%.1f merge %s %.2f ms e2e %.1f pk %.1f' % (c['value'], c['merge']['kind'], c['merge']['ms'], c['e2e']['value'], c.get('e2e_packed', {}).get('value', 0))))
PY

#!/bin/bash
# round 2, GPU call 12 (N GPUs): N=4: the driver's bench command; N=8: the hybrid merge at 30 / 50 %
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
if [ "$N" = "4" ]; then
  timeout 900 $TR --master-port 29811 bench.py --gpus 4 --steps 20 --warmup 3 > gpurun_out/bench_r2_n4.json 2> gpurun_out/bench_r2_n4.err; echo "bench n4 rc=$?"
  tail -c 500 gpurun_out/bench_r2_n4.err
else
  short="--gpus $N --steps 8 --warmup 3 --no-configs --no-e2e --no-cpu-baseline --no-job"
  for m in hybrid30 hybrid50; do
    timeout 300 $TR --master-port 29812 bench.py $short --merge $m > gpurun_out/bench_r2_n${N}_$m.json 2> gpurun_out/bench_r2_n${N}_$m.err; echo "$m rc=$?"
  done
fi
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_r2_n4.json') + glob.glob('gpurun_out/bench_r2_n8_hybrid*.json')):
    txt = [l for l in open(f).read().splitlines() if l.startswith('{')]
    if not txt:
        print(f, 'no json'); continue
    d = json.loads(txt[-1]); m = d.get('merge', {})
    print(f, 'value %.1f' % d.get('value', 0), 'merge', m.get('kind'), m.get('ms'), m.get('calibration_ms'), 'e2e', d.get('e2e', {}).get('value'), 'pk', d.get('e2e_packed', {}).get('value'), 'job', (d.get('job') or {}).get('gkmers_s'), 'check', (d.get('merge_check') or {}).get('merge_parity'))
    for n, c in (d.get('configs') or {}).items():
        print('   ', n, c.get('error') or ('%.1f merge %s %.2f ms' % (c['value'], c['merge']['kind'], c['merge']['ms'])))
PY

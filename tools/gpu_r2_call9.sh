#!/bin/bash
# round 2, GPU call 9 (N GPUs, default 8): hybrid merge (in-switch + peer memory side by side): correctness, calibration
N=${1:-8}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q -s > gpurun_out/pytest_multi9.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi9.log
grep -a "MULTI-GPU\|passed\|failed\|rc=\|Error\|error" gpurun_out/pytest_multi9.log | tail -6
short="--gpus $N --steps 8 --warmup 3 --no-configs --no-e2e --no-cpu-baseline"
timeout 400 $TR --master-port 29711 bench.py $short > gpurun_out/bench_r2_n${N}_auto.json 2> gpurun_out/bench_r2_n${N}_auto.err; echo "rc=$?"
tail -c 600 gpurun_out/bench_r2_n${N}_auto.err
python - <<PY
import json
txt = [l for l in open('gpurun_out/bench_r2_n${N}_auto.json').read().splitlines() if l.startswith('{')]
d = json.loads(txt[-1]); m = d['merge']
print('value %.1f merge %s %.2f ms calibration %s job %s merge_ms %s' % (d['value'], m['kind'], m['ms'], m.get('calibration_ms'), d['job']['gkmers_s'], d['job']['merge_ms']))
PY

#!/bin/bash
# round 2, GPU call 5 (2 GPUs): multi-GPU worker (parked k-mers across merges, in-switch merge), bench at N=2 with the
# merge inside the job, peer-memory merge beside it, knobs of the multimem kernel, reference arm under torchrun
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -15 gpurun_out/pytest_multi.log
timeout 900 $TR --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo "bench n2 rc=$?"
tail -c 1500 gpurun_out/bench_r2_n2.err
short="--gpus 2 --steps 8 --warmup 3 --no-configs --no-e2e --no-job --no-cpu-baseline"
timeout 300 $TR --master-port 29512 bench.py $short --merge peer > gpurun_out/bench_r2_n2_peer.json 2> gpurun_out/bench_r2_n2_peer.err
for o in "mm_unroll=1" "mm_unroll=2" "mm_unroll=8" "mm_grid=148" "mm_grid=296" "mm_grid=1184" "mm_unroll=8 --opt mm_grid=296"; do
  timeout 300 $TR --master-port 29513 bench.py $short --opt $o > "gpurun_out/bench_r2_n2_$(echo $o | tr ' =' '__' | tr -d '-').json" 2>> gpurun_out/bench_r2_n2_sweep.err
done
timeout 300 $TR --master-port 29514 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/bench_r2_n2_ref.json 2> gpurun_out/bench_r2_n2_ref.err; echo "ref n2 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/bench_r2_n2*.json')):
    try:
        d = json.load(open(f))
    except Exception as e:
        print(f, 'unreadable', e); continue
    m = d.get('merge', {})
    print(f, 'value %.2f' % d.get('value', 0), 'merge', m.get('kind'), m.get('ms'), m.get('multimem_unavailable'), 'e2e', d.get('e2e', {}).get('value'), 'e2e_packed', d.get('e2e_packed', {}).get('value'), 'cores', d.get('cpu_baseline', {}).get('cores'))
PY

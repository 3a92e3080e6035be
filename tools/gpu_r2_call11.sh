#!/bin/bash
# round 2, GPU call 11 (1 GPU): bulk-copy (TMA) copy-out of query pass 1: parity suite, fuzz, quick bench, source-level profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2d.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2d.log
tail -6 gpurun_out/pytest_gpu_r2d.log
timeout 300 python tools/gpu_fuzz.py 150 4242 > gpurun_out/fuzz_r2d.log 2>&1; tail -2 gpurun_out/fuzz_r2d.log
timeout 600 python bench.py --no-cpu-baseline --no-configs --no-job > gpurun_out/bench_r2d.json 2> gpurun_out/bench_r2d.err; echo "bench rc=$?"; tail -c 400 gpurun_out/bench_r2d.err
timeout 300 python tools/r2_sweep.py cfg2 "" > gpurun_out/sweep4_cfg2.jsonl 2>/dev/null
timeout 300 python tools/r2_sweep.py cfg3 "" > gpurun_out/sweep4_cfg3.jsonl 2>/dev/null
SKIP_BUILD=1 THRESHOLD=1 timeout 300 python tools/r2_sweep.py cfg4 "" > gpurun_out/sweep4_cfg4.jsonl 2>/dev/null
cat gpurun_out/sweep4_*.jsonl | cut -c1-400
args="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-job"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bin_kernel_sort -s 4 -c 1 -f -o gpurun_out/p1q2 python bench.py $args > gpurun_out/ncu_p1q2.log 2>&1
ncu -i gpurun_out/p1q2.ncu-rep --page source --csv --print-source sass > gpurun_out/p1q2_source.csv 2> /dev/null
ncu -i gpurun_out/p1q2.ncu-rep --page raw --csv > gpurun_out/p1q2_raw.csv 2> /dev/null
rm -f gpurun_out/p1q2.ncu-rep
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_r2d.json') if l.startswith('{')][-1])
print('value %.2f ins %.2f qry %.2f e2e %.2f packed %.2f' % (d['value'], d['insert_gkmers_s'], d['query_gkmers_s'], d['e2e']['value'], d.get('e2e_packed', {}).get('value', 0)))
PY

#!/bin/bash
# round 2, GPU call 6 (1 GPU): C++ host scenarios (per-k-mer loop rate), default bench line with the packed e2e leg,
# source-level ncu capture of query pass 1 (dynamic instruction mix), per-kernel times of the counting config, fuzz
mkdir -p gpurun_out
python -m pytest tests/test_cpp_host.py -m gpu -q > gpurun_out/pytest_cpp.log 2>&1
tests/cpp/test_host_classes /tmp > gpurun_out/cpp_host.log 2>&1; echo "cpp rc=$?"; tail -6 gpurun_out/cpp_host.log
timeout 900 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_r2b.err
args="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-job"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bin_kernel_sort -s 4 -c 1 -f -o gpurun_out/p1q python bench.py $args > gpurun_out/ncu_p1q.log 2>&1
ncu -i gpurun_out/p1q.ncu-rep --page source --csv --print-source sass > gpurun_out/p1q_source.csv 2> gpurun_out/p1q_source.err || ncu -i gpurun_out/p1q.ncu-rep --page source --csv > gpurun_out/p1q_source.csv 2>> gpurun_out/p1q_source.err
ncu -i gpurun_out/p1q.ncu-rep --page raw --csv > gpurun_out/p1q_raw.csv 2> /dev/null
rm -f gpurun_out/p1q.ncu-rep
ls -la gpurun_out/p1q_source.csv; head -c 600 gpurun_out/p1q_source.csv
# per-kernel durations of one counting-filter step
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 600 --csv \
  --log-file gpurun_out/launches_cfg4.csv python bench.py --config cfg4 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-job > gpurun_out/ncu_cfg4.log 2>&1
timeout 400 python tools/gpu_fuzz.py 240 20261018 > gpurun_out/fuzz_r2.log 2>&1; tail -3 gpurun_out/fuzz_r2.log
python - <<'PY'
import json
d = json.loads([l for l in open('gpurun_out/bench_r2b.json') if l.startswith('{')][-1])
print('value %.2f e2e %.2f packed %.2f sync %.2f' % (d['value'], d['e2e']['value'], d.get('e2e_packed', {}).get('value', 0), d.get('e2e_sync', {}).get('value', 0)))
for n, c in d['configs'].items():
    print(n, c.get('error') or ('%.2f ins %.2f qry %.2f e2e %.2f pk %.2f' % (c['value'], c['insert_gkmers_s'], c['query_gkmers_s'], c['e2e']['value'], c.get('e2e_packed', {}).get('value', 0))))
PY

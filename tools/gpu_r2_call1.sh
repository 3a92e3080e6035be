#!/bin/bash
# round 2, GPU call 1: tests, option sweeps (query overlap, prefetch), first bench line
mkdir -p gpurun_out
set -x
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
SKIP_BUILD=1 timeout 300 python tools/r2_sweep.py cfg2 "query_sub=1" "query_sub=1,query_probe_unroll=4" \
  "query_sub=2" "query_sub=4" "query_sub=8" "query_sub=4,query_p1_ctas=4" "query_sub=4,query_p1_ctas=2" \
  "query_sub=4,query_probe_unroll=2" "query_sub=4,query_probe_unroll=1" "query_sub=2,query_p1_ctas=4" \
  > gpurun_out/sweep_cfg2.jsonl 2> gpurun_out/sweep_cfg2.err
SKIP_BUILD=1 timeout 300 python tools/r2_sweep.py cfg3 "query_sub=1" "query_sub=4,query_p1_ctas=1" "query_sub=4,query_p1_ctas=2" \
  "query_sub=2,query_p1_ctas=2" > gpurun_out/sweep_cfg3.jsonl 2> gpurun_out/sweep_cfg3.err
timeout 400 python tools/r2_sweep.py cfg5b "bin_query_mode=-1" "bin_query_mode=1,query_sub=1" "bin_query_mode=1,query_sub=1,bin_prefetch=1" \
  "bin_query_mode=1,query_sub=1,bin_max_parts=1024" "bin_query_mode=1,query_sub=4" "bin_prefetch=1" \
  > gpurun_out/sweep_cfg5b.jsonl 2> gpurun_out/sweep_cfg5b.err
SKIP_BUILD=1 THRESHOLD=1 timeout 400 python tools/r2_sweep.py cfg4 "bin_query_mode=-1" "bin_query_mode=1,query_sub=1" \
  "bin_query_mode=1,query_sub=1,bin_prefetch=1" "bin_query_mode=1,query_sub=1,bin_max_parts=1024" "bin_query_mode=1,query_sub=4" \
  > gpurun_out/sweep_cfg4.jsonl 2> gpurun_out/sweep_cfg4.err
timeout 600 python bench.py > gpurun_out/bench_r2_a.json 2> gpurun_out/bench_r2_a.err; echo "bench rc=$?"
tail -c 1500 gpurun_out/bench_r2_a.err
cat gpurun_out/sweep_cfg2.jsonl | cut -c1-300

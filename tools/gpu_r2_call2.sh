#!/bin/bash
# round 2, GPU call 2: counting partitioned query fix, L1-split hypothesis of the probe kernel, probe load flavours
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_production_geometry.py -m gpu -q > gpurun_out/pytest_gpu2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu2.log
tail -5 gpurun_out/pytest_gpu2.log
SKIP_BUILD=1 timeout 300 python tools/r2_sweep.py cfg2 "query_sub=1" "query_sub=1,probe_carveout=1" \
  "query_sub=1,probe_carveout=1,probe_ld=1" "query_sub=1,probe_carveout=1,probe_ld=2" "query_sub=1,probe_ld=1" "query_sub=1,probe_ld=2" \
  "query_sub=1,query_probe_unroll=2" "query_sub=1,query_probe_unroll=4" "query_sub=4" "query_sub=4,probe_ld=1" "query_sub=4,probe_ld=2" \
  "query_sub=2,probe_ld=1" "query_sub=4,probe_ld=1,query_p1_ctas=4" "query_sub=4,probe_ld=1,query_p1_ctas=2" \
  > gpurun_out/sweep2_cfg2.jsonl 2> gpurun_out/sweep2_cfg2.err
SKIP_BUILD=1 timeout 300 python tools/r2_sweep.py cfg3 "query_sub=1" > gpurun_out/sweep2_cfg3.jsonl 2> gpurun_out/sweep2_cfg3.err
timeout 400 python tools/r2_sweep.py cfg5b "bin_query_mode=-1" "bin_query_mode=1,query_sub=1" "bin_query_mode=1,query_sub=1,bin_prefetch=1" \
  "bin_query_mode=1,query_sub=1,bin_prefetch=1,query_probe_unroll=4" "bin_prefetch=1" \
  > gpurun_out/sweep2_cfg5b.jsonl 2> gpurun_out/sweep2_cfg5b.err
SKIP_BUILD=1 THRESHOLD=1 timeout 400 python tools/r2_sweep.py cfg4 "bin_query_mode=-1" "bin_query_mode=1,query_sub=1" \
  "bin_query_mode=1,query_sub=1,bin_prefetch=1" "bin_query_mode=1,query_sub=1,bin_prefetch=1,query_probe_unroll=4" \
  > gpurun_out/sweep2_cfg4.jsonl 2> gpurun_out/sweep2_cfg4.err
cat gpurun_out/sweep2_cfg2.jsonl | cut -c1-260

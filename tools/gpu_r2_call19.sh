#!/bin/bash
# round 2, GPU call 19 (1 GPU): sort-bin kernel in 384-thread CTAs (three per SM, 36 warps, 56 registers) against the defaults
mkdir -p gpurun_out
timeout 300 python tools/r2_sweep.py cfg2 "" "bin_threads=384" "bin_threads=512" > gpurun_out/sweep7_cfg2.jsonl 2>gpurun_out/sweep7.err
timeout 300 python tools/r2_sweep.py cfg3 "" "bin_threads=384" "bin_threads=256" > gpurun_out/sweep7_cfg3.jsonl 2>>gpurun_out/sweep7.err
SKIP_BUILD=1 THRESHOLD=1 timeout 300 python tools/r2_sweep.py cfg4 "" "bin_threads=384" > gpurun_out/sweep7_cfg4.jsonl 2>>gpurun_out/sweep7.err
cat gpurun_out/sweep7_*.jsonl | cut -c1-330; tail -3 gpurun_out/sweep7.err

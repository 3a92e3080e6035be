#!/bin/bash
# round 2, GPU call 16 (1 GPU): pass-1 shape experiment (rounds of 2 windows, five 256-thread CTAs per SM, 48 registers)
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_production_geometry.py tests/test_gpu_parity.py -m gpu -q -x -k "binned or partitioned or full_size" > gpurun_out/pytest_gpu_r2h.log 2>&1; tail -2 gpurun_out/pytest_gpu_r2h.log
timeout 300 python tools/r2_sweep.py cfg2 "" > gpurun_out/sweep6_cfg2.jsonl 2>/dev/null
timeout 300 python tools/r2_sweep.py cfg3 "" > gpurun_out/sweep6_cfg3.jsonl 2>/dev/null
SKIP_BUILD=1 THRESHOLD=1 timeout 300 python tools/r2_sweep.py cfg4 "" > gpurun_out/sweep6_cfg4.jsonl 2>/dev/null
BUILD_REPS=16 timeout 300 python tools/r2_sweep.py cfg5b "" > gpurun_out/sweep6_cfg5b.jsonl 2>/dev/null
cat gpurun_out/sweep6_*.jsonl | cut -c1-330

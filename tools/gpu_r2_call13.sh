#!/bin/bash
# round 2, GPU call 13 (1 GPU): spaced-seed query with all probes of a window in flight: parity + cfg5a / cfg5b rates
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x -k "spaced or golden or cfg5 or fuzz or packed" > gpurun_out/pytest_gpu_r2e.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2e.log
tail -4 gpurun_out/pytest_gpu_r2e.log
timeout 200 python tools/gpu_fuzz.py 90 99 > gpurun_out/fuzz_r2e.log 2>&1; tail -2 gpurun_out/fuzz_r2e.log
timeout 300 python tools/r2_sweep.py cfg5a "" "query_mode=1" > gpurun_out/sweep5_cfg5a.jsonl 2>/dev/null
BUILD_REPS=16 timeout 300 python tools/r2_sweep.py cfg5b "" > gpurun_out/sweep5_cfg5b.jsonl 2>/dev/null
cat gpurun_out/sweep5_*.jsonl | cut -c1-330

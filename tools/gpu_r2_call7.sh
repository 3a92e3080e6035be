#!/bin/bash
# round 2, GPU call 7 (1 GPU): full GPU suite on the new seeding / per-thread queues, C++ per-k-mer loop rate, evidence round
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2c.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2c.log
tail -6 gpurun_out/pytest_gpu_r2c.log
tests/cpp/test_host_classes /tmp > gpurun_out/cpp_host.log 2>&1; echo "cpp rc=$?"; grep -i "threaded" gpurun_out/cpp_host.log
timeout 300 python tools/gpu_fuzz.py 120 777 > gpurun_out/fuzz_r2c.log 2>&1; tail -2 gpurun_out/fuzz_r2c.log
timeout 1500 bash tools/gpu_profile_round.sh r2c

#!/bin/bash
# round 2, GPU call 4: the whole GPU suite (packed input tests included), then the round's evidence (bench line,
# reference arm, ncu launch list + full capture of a short run of the same command)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_r2.log
tail -8 gpurun_out/pytest_gpu_r2.log
timeout 1500 bash tools/gpu_profile_round.sh r2a

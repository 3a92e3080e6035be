#!/bin/bash
# round 2, final evidence on one GPU: GPU suite, smoke(), the default bench line + reference arm + ncu launch list + full
# capture (tools/gpu_profile_round.sh), source-level capture of query pass 1 of the final kernels
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu_final.log
tail -4 gpurun_out/pytest_gpu_final.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke_final.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke_final.log
timeout 1500 bash tools/gpu_profile_round.sh r2f
args="--steps 2 --warmup 1 --no-e2e --no-cpu-baseline --no-configs --no-job"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:bin_kernel_sort -s 4 -c 1 -f -o gpurun_out/p1q3 python bench.py $args > gpurun_out/ncu_p1q3.log 2>&1
ncu -i gpurun_out/p1q3.ncu-rep --page source --csv --print-source sass > gpurun_out/p1q3_source.csv 2> /dev/null
rm -f gpurun_out/p1q3.ncu-rep
ls -la gpurun_out/p1q3_source.csv

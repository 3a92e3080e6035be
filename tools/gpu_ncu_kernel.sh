#!/bin/bash
# usage: tools/gpu_ncu_kernel.sh <kernel-base-name> <skip> <count> <out-name> [bench args]
#   one ncu --set full capture (with source) of <count> launches of one kernel after skipping <skip> of them
k=$1; skip=$2; cnt=$3; out=$4; shift 4
python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 1 "$@" > gpurun_out/ncuk_plain.log 2>&1 || { echo FAIL; tail -3 gpurun_out/ncuk_plain.log; exit 1; }
ncu --set full --import-source on --clock-control none -k "$k" -s $skip -c $cnt -f -o gpurun_out/$out python bench.py --no-e2e --no-cpu-baseline --steps 1 --warmup 1 "$@" > gpurun_out/ncuk.log 2>&1
ls -la gpurun_out/$out.ncu-rep

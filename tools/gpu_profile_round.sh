#!/bin/bash
# usage: tools/gpu_profile_round.sh <tag>   (on the GPU box)
#   1. bench.py with its defaults                     -> gpurun_out/bench_<tag>.json
#   2. ncu launch list of a short run of the same command (gpu__time_duration + DRAM bytes per launch)
#   3. one ncu --set full capture of the four kernels of a step (pass 1 build, pass 2 build, pass 1 query, pass 2 query)
tag=$1
python -c "import bench; print(bench.source_sha())" > gpurun_out/source_sha_$tag.txt
python bench.py > gpurun_out/bench_$tag.json 2> gpurun_out/bench_$tag.err || { echo "bench failed"; tail -5 gpurun_out/bench_$tag.err; exit 1; }
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref_$tag.json 2>> gpurun_out/bench_$tag.err
args="--steps 8 --warmup 3 --no-e2e --no-cpu-baseline --no-configs --no-job"
python bench.py $args > gpurun_out/prof_plain_$tag.json 2>> gpurun_out/bench_$tag.err || { echo "short bench failed"; exit 1; }
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches_$tag.csv python bench.py $args > gpurun_out/ncu_list_$tag.log 2>&1
# full capture of the matching launches 18..25 = last 3 build pass-1 launches, both pass-2 launches of the timed
# region, and the first query (pass 1, pass 2, finalize); 13 matching launches belong to the warm-up
ncu --set full --clock-control none -k regex:"bin_kernel_sort|apply_bins|probe_bins|finalize_hits" \
    -s 18 -c 8 -f -o gpurun_out/prof_$tag python bench.py $args > gpurun_out/ncu_full_$tag.log 2>&1
# the report itself is too large to travel back (gpurun_out is limited to 64 MiB): keep its raw page as CSV
ncu -i gpurun_out/prof_$tag.ncu-rep --page raw --csv > gpurun_out/prof_${tag}_raw.csv 2> /dev/null
rm -f gpurun_out/prof_$tag.ncu-rep
ls -la gpurun_out/prof_${tag}_raw.csv gpurun_out/launches_$tag.csv
python -c "
import json
d=json.load(open('gpurun_out/bench_$tag.json'))
print('value %.2f insert %.2f query %.2f e2e %.2f launches %d' % (d['value'], d['insert_gkmers_s'], d['query_gkmers_s'], d['e2e']['value'], d['gpu_launches']))
print('roofline (query) frac %.3f build frac %.3f step frac %.3f' % (d['roofline']['frac'], d['roofline_build']['frac'], d['roofline_step']['frac']))
print('cpu', d['cpu_baseline']['value'], d['cpu_baseline']['cores'])
r=json.load(open('gpurun_out/bench_ref_$tag.json')); print('ref arm', r['value'], r['cpu_baseline']['cores'])
"

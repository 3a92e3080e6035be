#!/bin/bash
# round 2, GPU call 3 (2 GPUs): multi-GPU test, bench at N=2 (merge inside the timed job), reference arm under torchrun
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q > gpurun_out/pytest_multi.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_multi.log
tail -5 gpurun_out/pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 3 \
  > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; echo "bench n2 rc=$?"
tail -c 2000 gpurun_out/bench_r2_n2.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 \
  > gpurun_out/bench_r2_n2_ref.json 2> gpurun_out/bench_r2_n2_ref.err; echo "ref n2 rc=$?"
cut -c1-400 gpurun_out/bench_r2_n2_ref.json

#!/bin/bash
# 8-GPU pass (gpurun --gpus 8): the default workload under torchrun, one rank per GPU
tag=${1:-n8}
set -x
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 8 --steps 3 --warmup 3 --no-cpu --no-eager --no-shares > gpurun_out/bench_$tag.log 2>&1; grep '{"metric' gpurun_out/bench_$tag.log | cut -c1-1200

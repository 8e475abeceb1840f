#!/bin/bash
# 2-GPU pass (gpurun --gpus 2): real-NCCL DD-PPO correctness test + bench under torchrun (savi and interactive configs)
tag=${1:-n2}
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddppo_nccl.py -m gpu -x -q > gpurun_out/tests_$tag.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_$tag.log
tail -4 gpurun_out/tests_$tag.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-eager --no-shares > gpurun_out/bench_$tag.log 2>&1; grep '{"metric' gpurun_out/bench_$tag.log | cut -c1-600
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config interactive --steps 2 --warmup 3 --no-cpu --no-shares > gpurun_out/bench_${tag}_interactive.log 2>&1; grep '{"metric' gpurun_out/bench_${tag}_interactive.log | cut -c1-600

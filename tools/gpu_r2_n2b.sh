#!/bin/bash
# 2-GPU pass with the final build of the round: real-NCCL DD-PPO correctness test + bench under torchrun
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddppo_nccl.py -m gpu -x -q > gpurun_out/r02_tests_n2b.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_n2b.log
tail -4 gpurun_out/r02_tests_n2b.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-eager --no-shares > gpurun_out/r02_bench_n2b.log 2>&1; grep '{"metric' gpurun_out/r02_bench_n2b.log | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config interactive --steps 2 --warmup 3 --no-cpu --no-shares > gpurun_out/r02_bench_n2b_interactive.log 2>&1; grep '{"metric' gpurun_out/r02_bench_n2b_interactive.log | cut -c1-600
timeout 600 python tools/wgrad_conv_bench.py 4800 > gpurun_out/r02_wgrad_conv_bench_n2b.txt 2>&1; grep -E "wgrad_ms|gn_bwd" gpurun_out/r02_wgrad_conv_bench_n2b.txt | cut -c1-200

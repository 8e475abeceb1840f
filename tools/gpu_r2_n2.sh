#!/bin/bash
# 2-GPU pass: real-NCCL DD-PPO correctness test + bench under torchrun (savi and interactive configs)
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_ddppo_nccl.py -m gpu -x -q > gpurun_out/r02_tests_n2.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_n2.log
tail -8 gpurun_out/r02_tests_n2.log | cut -c1-200
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-eager --no-shares > gpurun_out/r02_bench_n2.log 2>&1; tail -1 gpurun_out/r02_bench_n2.log | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --config interactive --steps 2 --warmup 2 --no-cpu --no-shares > gpurun_out/r02_bench_n2_interactive.log 2>&1; tail -1 gpurun_out/r02_bench_n2_interactive.log | cut -c1-1200
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --config audio_sweep --no-sweep --no-cpu > gpurun_out/r02_bench_n2_audio.log 2>&1; tail -1 gpurun_out/r02_bench_n2_audio.log | cut -c1-600

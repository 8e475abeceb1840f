#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_graphs.py tests/test_gpu_obs.py tests/test_gpu_policy.py -m gpu -x -q > gpurun_out/r02_tests_bb.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_bb.log
tail -30 gpurun_out/r02_tests_bb.log | cut -c1-250
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares --no-e2e --regime trainable > gpurun_out/r02_bench_bb_trainable.log 2>&1; grep '{"metric' gpurun_out/r02_bench_bb_trainable.log | cut -c1-900

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_bwd.py tests/test_gpu_avnav.py tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r02_tests_b.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_b.log
tail -25 gpurun_out/r02_tests_b.log
timeout 600 python bench.py --steps 2 --warmup 1 --regime trainable --no-e2e --no-cpu --no-eager --no-shares > gpurun_out/r02_bench_b_trainable.log 2>&1; tail -2 gpurun_out/r02_bench_b_trainable.log
AVL_REGIME=trainable timeout 600 python tools/profile_step.py 150 > gpurun_out/r02_profile_trainable_b.txt 2>&1
grep -A22 "PPO update" gpurun_out/r02_profile_trainable_b.txt | cut -c1-150 | head -30

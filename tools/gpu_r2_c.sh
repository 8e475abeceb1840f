#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc_bwd.py tests/test_gpu_avnav.py tests/test_gpu_nn.py -m gpu -x -q > gpurun_out/r02_tests_c.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_c.log
tail -25 gpurun_out/r02_tests_c.log
timeout 600 python tools/wgrad_conv_bench.py 4800 > gpurun_out/r02_wgrad_conv_bench_c.txt 2>&1; cat gpurun_out/r02_wgrad_conv_bench_c.txt | grep layer
timeout 600 python bench.py --steps 2 --warmup 1 --regime trainable --no-e2e --no-cpu --no-eager --no-shares > gpurun_out/r02_bench_c_trainable.log 2>&1; tail -1 gpurun_out/r02_bench_c_trainable.log | cut -c1-900
AVL_REGIME=trainable timeout 600 python tools/profile_step.py 150 > gpurun_out/r02_profile_trainable_c.txt 2>&1
grep -A22 "PPO update" gpurun_out/r02_profile_trainable_c.txt | cut -c1-150 | head -30

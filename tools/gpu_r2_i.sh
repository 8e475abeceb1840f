#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_graphs.py tests/test_gpu_golden.py tests/test_gpu_policy.py -m gpu -x -q > gpurun_out/r02_tests_i.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_i.log
tail -40 gpurun_out/r02_tests_i.log | cut -c1-220
timeout 900 python bench.py --steps 3 --warmup 3 --regime frozen --no-cpu --no-eager --no-shares > gpurun_out/r02_bench_i_frozen.log 2>&1; tail -3 gpurun_out/r02_bench_i_frozen.log | cut -c1-900

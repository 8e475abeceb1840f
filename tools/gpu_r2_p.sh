#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_graphs.py tests/test_gpu_obs.py tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r02_tests_p.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_p.log
tail -25 gpurun_out/r02_tests_p.log | cut -c1-250
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares --regime frozen > gpurun_out/r02_bench_p.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_p.log | cut -c1-1100; tail -5 gpurun_out/r02_bench_p.log | cut -c1-300

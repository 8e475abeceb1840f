#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_interactive.py tests/test_gpu_policy.py tests/test_gpu_dialog.py -m gpu -x -q > gpurun_out/r02_tests_k.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_k.log
tail -6 gpurun_out/r02_tests_k.log | cut -c1-220
timeout 900 python bench.py --config interactive --steps 2 --warmup 2 --no-cpu --no-shares > gpurun_out/r02_bench_k_interactive.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_k_interactive.log | cut -c1-700
timeout 900 python bench.py --config distractor --steps 2 --warmup 2 --no-cpu --no-shares --no-e2e > gpurun_out/r02_bench_k_distractor.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_k_distractor.log | cut -c1-700

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_graphs.py tests/test_gpu_interactive.py tests/test_gpu_dialog.py -m gpu -x -q > gpurun_out/r02_tests_l.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_l.log
tail -30 gpurun_out/r02_tests_l.log | cut -c1-220
timeout 900 python bench.py --config interactive --steps 2 --warmup 3 --no-cpu --no-shares > gpurun_out/r02_bench_l_interactive.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_l_interactive.log | cut -c1-700

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_dialog.py tests/test_gpu_interactive.py tests/test_gpu_step_graphs.py tests/test_gpu_tc.py -m gpu -x -q > gpurun_out/r02_tests_dd.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_dd.log
tail -5 gpurun_out/r02_tests_dd.log | cut -c1-250
timeout 900 python bench.py --config interactive --steps 2 --warmup 3 --no-cpu --no-shares > gpurun_out/r02_bench_dd_interactive.log 2>&1; grep '{"metric' gpurun_out/r02_bench_dd_interactive.log | cut -c1-900
AVL_POLICY=interactive timeout 900 python tools/profile_step.py 150 > gpurun_out/r02_profile_interactive_dd.txt 2>&1; grep -A12 "10 rollout steps" gpurun_out/r02_profile_interactive_dd.txt | cut -c1-170

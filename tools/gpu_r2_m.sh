#!/bin/bash
# full GPU suite + the default bench line exactly as the driver runs it + per-kernel profile of both regimes
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_m.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_m.log
tail -6 gpurun_out/r02_tests_m.log | cut -c1-220
( time timeout 1200 python bench.py ) > gpurun_out/r02_bench_m.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_m.log | cut -c1-3000; tail -4 gpurun_out/r02_bench_m.log
timeout 600 python tools/profile_step.py 150 > gpurun_out/r02_profile_frozen_m.txt 2>&1; tail -3 gpurun_out/r02_profile_frozen_m.txt | cut -c1-200
AVL_REGIME=trainable timeout 600 python tools/profile_step.py 150 > gpurun_out/r02_profile_trainable_m.txt 2>&1; tail -3 gpurun_out/r02_profile_trainable_m.txt | cut -c1-200

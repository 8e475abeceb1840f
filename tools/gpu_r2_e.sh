#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_e.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_e.log
tail -8 gpurun_out/r02_tests_e.log | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_e.log 2>&1; tail -1 gpurun_out/r02_bench_e.log | cut -c1-1200
timeout 600 python tools/profile_step.py 150 > gpurun_out/r02_profile_frozen_e.txt 2>&1
grep -A14 "10 rollout steps (64" gpurun_out/r02_profile_frozen_e.txt | cut -c1-150; grep -A16 "PPO update" gpurun_out/r02_profile_frozen_e.txt | cut -c1-150

#!/bin/bash
# round 2, first GPU pass: new tests, bench in both regimes, CUPTI profile of the trainable update
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_a.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_a.log
tail -15 gpurun_out/r02_tests_a.log
timeout 900 python bench.py --steps 3 --warmup 3 > gpurun_out/r02_bench_a.log 2>&1; echo "bench rc=$?"
tail -3 gpurun_out/r02_bench_a.log
AVL_REGIME=trainable timeout 600 python tools/profile_step.py 150 > gpurun_out/r02_profile_trainable_a.txt 2>&1
grep -A22 "PPO update" gpurun_out/r02_profile_trainable_a.txt | head -40
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r02_bench_ref_a.log 2>&1; tail -1 gpurun_out/r02_bench_ref_a.log

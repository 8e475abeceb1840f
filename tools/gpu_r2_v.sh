#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1500 python bench.py ) > gpurun_out/r02_bench_v.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_v.log | cut -c1-200; tail -4 gpurun_out/r02_bench_v.log
timeout 600 python tools/profile_step.py 150 > gpurun_out/r02_profile_frozen_v.txt 2>&1; tail -3 gpurun_out/r02_profile_frozen_v.txt | cut -c1-200
AVL_REGIME=trainable timeout 600 python tools/profile_step.py 150 > gpurun_out/r02_profile_trainable_v.txt 2>&1; tail -3 gpurun_out/r02_profile_trainable_v.txt | cut -c1-200
timeout 900 python bench.py --config interactive --steps 2 --warmup 3 --no-cpu --no-shares > gpurun_out/r02_bench_v_interactive.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_v_interactive.log | cut -c1-400

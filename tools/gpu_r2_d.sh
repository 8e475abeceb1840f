#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_graph_env.py tests/test_gpu_interactive.py tests/test_gpu_tc_bwd.py tests/test_gpu_obs.py -m gpu -x -q > gpurun_out/r02_tests_d.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_d.log
tail -30 gpurun_out/r02_tests_d.log | cut -c1-200
timeout 600 python -m pytest tests/test_gpu_golden.py -m gpu -x -q -k "interactive or audiogoal" > gpurun_out/r02_tests_d2.log 2>&1; tail -5 gpurun_out/r02_tests_d2.log
timeout 600 python tools/wgrad_conv_bench.py 4800 > gpurun_out/r02_wgrad_conv_bench_d.txt 2>&1; grep wgrad_ms gpurun_out/r02_wgrad_conv_bench_d.txt | cut -c1-200
timeout 600 python bench.py --steps 2 --warmup 1 --regime trainable --no-e2e --no-cpu --no-eager --no-shares > gpurun_out/r02_bench_d_trainable.log 2>&1; tail -1 gpurun_out/r02_bench_d_trainable.log | cut -c1-800
timeout 900 python bench.py --config interactive --steps 2 --warmup 1 --no-cpu > gpurun_out/r02_bench_d_interactive.log 2>&1; tail -3 gpurun_out/r02_bench_d_interactive.log | cut -c1-1500

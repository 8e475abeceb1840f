#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_j.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_j.log
tail -12 gpurun_out/r02_tests_j.log | cut -c1-220
timeout 600 python tools/wgrad_conv_bench.py 4800 > gpurun_out/r02_wgrad_conv_bench_j.txt 2>&1; grep -E "wgrad_ms|gn_bwd" gpurun_out/r02_wgrad_conv_bench_j.txt | cut -c1-170
timeout 900 python bench.py --config interactive --steps 2 --warmup 2 --no-cpu --no-shares > gpurun_out/r02_bench_j_interactive.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_j_interactive.log | cut -c1-900
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares > gpurun_out/r02_bench_j.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_j.log | cut -c1-1200

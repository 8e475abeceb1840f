#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_avnav.py tests/test_gpu_nn.py -m gpu -x -q > gpurun_out/r02_tests_aa.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_aa.log
tail -25 gpurun_out/r02_tests_aa.log | cut -c1-250
timeout 600 python tools/wgrad_conv_bench.py 4800 > gpurun_out/r02_wgrad_conv_bench_aa.txt 2>&1; grep -E "wgrad_ms" gpurun_out/r02_wgrad_conv_bench_aa.txt | cut -c1-200
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares --no-e2e --regime trainable > gpurun_out/r02_bench_aa_trainable.log 2>&1; grep '{"metric' gpurun_out/r02_bench_aa_trainable.log | cut -c1-900

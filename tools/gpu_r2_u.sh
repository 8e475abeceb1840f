#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_nn.py tests/test_gpu_golden.py tests/test_gpu_avnav.py -m gpu -x -q > gpurun_out/r02_tests_u.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_u.log
tail -30 gpurun_out/r02_tests_u.log | cut -c1-250
timeout 300 python tools/small_batch_conv_probe.py 64 > gpurun_out/r02_small_batch_conv_probe_u.txt 2>&1; grep "^B=" gpurun_out/r02_small_batch_conv_probe_u.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares --regime frozen > gpurun_out/r02_bench_u.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_u.log | cut -c1-1300

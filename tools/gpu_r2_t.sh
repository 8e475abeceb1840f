#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "tma_im2col or tc_conv" > gpurun_out/r02_tests_t.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_t.log
tail -30 gpurun_out/r02_tests_t.log | cut -c1-250
timeout 600 python tools/tma_conv_bench.py 4800 > gpurun_out/r02_tma_conv_bench_t.txt 2>&1; grep "^B=" gpurun_out/r02_tma_conv_bench_t.txt

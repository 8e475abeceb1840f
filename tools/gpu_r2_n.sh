#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_nn.py tests/test_gpu_policy.py tests/test_gpu_golden.py -m gpu -x -q > gpurun_out/r02_tests_n.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_n.log
tail -8 gpurun_out/r02_tests_n.log | cut -c1-220
for m in 0 1; do for b in 4800 64; do AVL_ATTN_TC=$m timeout 300 python tools/attn_bench.py $b; done; done > gpurun_out/r02_attn_bench_n.txt 2>&1
cat gpurun_out/r02_attn_bench_n.txt | grep -v Warn
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares --regime frozen > gpurun_out/r02_bench_n.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_n.log | cut -c1-900

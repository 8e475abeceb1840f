#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_nn.py -m gpu -x -q > gpurun_out/r02_tests_x.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_x.log
tail -12 gpurun_out/r02_tests_x.log | cut -c1-250
for t in 0 1; do AVL_HALO_TMA=$t timeout 300 python tools/halo_f16_bench.py 4800; AVL_HALO_TMA=$t timeout 300 python tools/halo_f16_bench.py 64; done > gpurun_out/r02_halo_tma_bench_x.txt 2>&1
grep -E "^--|^B=.*(layer|conv1)" gpurun_out/r02_halo_tma_bench_x.txt

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_obs.py -m gpu -x -q -k "attention or obs" > gpurun_out/r02_tests_o.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_o.log
tail -4 gpurun_out/r02_tests_o.log | cut -c1-220
AVL_HOST_BUFFERS=1 timeout 300 python tools/host_profile.py > gpurun_out/r02_host_profile_o_e2e.txt 2>&1; grep -E "wall per|function calls" gpurun_out/r02_host_profile_o_e2e.txt; grep -A45 "Ordered by: cumulative" gpurun_out/r02_host_profile_o_e2e.txt | cut -c1-160 | head -60
timeout 300 python tools/small_batch_conv_probe.py 64 > gpurun_out/r02_small_batch_conv_probe_o.txt 2>&1; grep "^B=" gpurun_out/r02_small_batch_conv_probe_o.txt
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares --regime frozen > gpurun_out/r02_bench_o.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_o.log | cut -c1-1100

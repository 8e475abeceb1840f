#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_step_graphs.py tests/test_gpu_obs.py -m gpu -x -q > gpurun_out/r02_tests_s.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_s.log
tail -3 gpurun_out/r02_tests_s.log | cut -c1-250
timeout 600 python tools/e2e_step_probe.py > gpurun_out/r02_e2e_step_probe_s.txt 2>&1; grep -E "device time|step span|launch|wait|stage|sum|streaming|native" gpurun_out/r02_e2e_step_probe_s.txt | cut -c1-150
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares --regime frozen > gpurun_out/r02_bench_s.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_s.log | cut -c1-1300

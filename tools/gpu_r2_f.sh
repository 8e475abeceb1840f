#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_golden.py tests/test_gpu_policy.py tests/test_gpu_obs.py tests/test_gpu_interactive.py -m gpu -x -q > gpurun_out/r02_tests_f.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_f.log
tail -6 gpurun_out/r02_tests_f.log | cut -c1-200
timeout 300 python tools/host_time.py > gpurun_out/r02_host_time_f.txt 2>&1; tail -4 gpurun_out/r02_host_time_f.txt
timeout 300 python tools/host_profile.py > gpurun_out/r02_host_profile_f.txt 2>&1; grep -A45 "Ordered by: internal time" gpurun_out/r02_host_profile_f.txt | cut -c1-150 | head -50
timeout 600 python bench.py --config audio_sweep > gpurun_out/r02_bench_f_audio.log 2>&1; tail -1 gpurun_out/r02_bench_f_audio.log | cut -c1-3000
timeout 600 python bench.py --config avnav --steps 2 --warmup 1 > gpurun_out/r02_bench_f_avnav.log 2>&1; tail -1 gpurun_out/r02_bench_f_avnav.log | cut -c1-1200
timeout 600 python bench.py --steps 2 --warmup 2 --regime frozen --no-cpu --no-eager --no-shares > gpurun_out/r02_bench_f_frozen.log 2>&1; tail -1 gpurun_out/r02_bench_f_frozen.log | cut -c1-900

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares --regime frozen > gpurun_out/r02_bench_q.log 2>&1; grep '^{"metric' gpurun_out/r02_bench_q.log | cut -c1-1300
AVL_HOST_BUFFERS=1 timeout 300 python tools/host_profile.py > gpurun_out/r02_host_profile_q_e2e.txt 2>&1; grep -E "wall per" gpurun_out/r02_host_profile_q_e2e.txt

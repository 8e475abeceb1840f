#!/bin/bash
# cluster split-K + narrow 3x tiles: parity, bench, stream trace, host profile
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -x -q -m gpu > gpurun_out/pytest_ck_tc.log 2>&1; echo "tc pytest rc=$?"; tail -3 gpurun_out/pytest_ck_tc.log
timeout 900 python -m pytest tests -x -q -m gpu > gpurun_out/pytest_gpu30.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu30.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v20.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r01_v20.log | cut -c1-400
AVL_TRACE=gpurun_out/trace_rollout_v20.json timeout 600 python tools/profile_step.py 150 1 > gpurun_out/profile_step_v20.log 2>&1; echo "profile rc=$?"
gzip -f gpurun_out/trace_rollout_v20.json
timeout 300 python tools/host_profile.py > gpurun_out/host_profile_v20.log 2>&1; echo "host profile rc=$?"
timeout 300 python tools/x3_gemm_bench.py > gpurun_out/x3_bench_v20.log 2>&1; tail -12 gpurun_out/x3_bench_v20.log

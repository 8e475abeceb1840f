#!/bin/bash
# prefetch (rollout + update), channel-split audio, cross-attention loads, fused env step: parity, bench, trace
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu41.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu32.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v31.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r01_v22.log | cut -c1-1200
AVL_TRACE=gpurun_out/trace_rollout_v31.json AVL_TRACE_UPDATE=gpurun_out/trace_update_v31.json timeout 600 python tools/profile_step.py 150 1 > gpurun_out/profile_step_v31.log 2>&1; echo "profile rc=$?"
gzip -f gpurun_out/trace_rollout_v31.json gpurun_out/trace_update_v31.json
timeout 300 python tools/host_time.py > gpurun_out/host_time_v31.log 2>&1; tail -4 gpurun_out/host_time_v31.log

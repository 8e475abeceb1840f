#!/bin/bash
# One GPU-box pass: the whole `-m gpu` suite, the default bench line (both regimes, both baselines, roofline) and the
# per-kernel profiles.  Run under gpurun:  gpurun --timeout 2400 -- 'bash tools/gpu_check.sh <tag>'
tag=${1:-check}
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/tests_$tag.log 2>&1; echo "tests rc=$?" >> gpurun_out/tests_$tag.log
tail -5 gpurun_out/tests_$tag.log | cut -c1-250
( time timeout 1500 python bench.py ) > gpurun_out/bench_$tag.log 2>&1; grep '{"metric' gpurun_out/bench_$tag.log | cut -c1-300
timeout 900 python bench.py --config interactive --steps 2 --warmup 3 --no-cpu --no-shares > gpurun_out/bench_${tag}_interactive.log 2>&1; grep '{"metric' gpurun_out/bench_${tag}_interactive.log | cut -c1-300
timeout 600 python tools/profile_step.py 150 > gpurun_out/profile_frozen_$tag.txt 2>&1
AVL_REGIME=trainable timeout 600 python tools/profile_step.py 150 > gpurun_out/profile_trainable_$tag.txt 2>&1

#!/bin/bash
set -x
mkdir -p gpurun_out
( time timeout 1500 python bench.py ) > gpurun_out/bench_r02_final2.log 2>&1; grep '{"metric' gpurun_out/bench_r02_final2.log | cut -c1-200
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/bench_r02_final2_reference.log 2>&1; grep '{"' gpurun_out/bench_r02_final2_reference.log | cut -c1-600

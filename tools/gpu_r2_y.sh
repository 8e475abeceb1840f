#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r02_tests_y.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_y.log
tail -5 gpurun_out/r02_tests_y.log | cut -c1-250
( time timeout 1500 python bench.py ) > gpurun_out/r02_bench_y.log 2>&1; grep '{"metric' gpurun_out/r02_bench_y.log | cut -c1-300; tail -4 gpurun_out/r02_bench_y.log

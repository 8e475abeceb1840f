#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_audio.py -m gpu -x -q > gpurun_out/r02_tests_ee.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_ee.log
tail -30 gpurun_out/r02_tests_ee.log | cut -c1-250

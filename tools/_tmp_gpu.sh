#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_audio.py -x -q > gpurun_out/tests_audio_spectral.log 2>&1; tail -15 gpurun_out/tests_audio_spectral.log | cut -c1-250
timeout 900 python bench.py --config audio_sweep --no-cpu > gpurun_out/bench_audio_spectral.log 2>&1; tail -3 gpurun_out/bench_audio_spectral.log | cut -c1-3000

#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_audio.py -x -q 2>&1 | tail -2
timeout 600 python bench.py --config audio_sweep --no-cpu > gpurun_out/bench_audio_b4.log 2>&1; grep '{"metric' gpurun_out/bench_audio_b4.log | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('audio', d['value'], d['roofline']['frac'], [ (p['n_envs'], p['ms'], p['env_steps_per_s'], p['hbm_frac']) for p in d['spectral_banks']['points']])"

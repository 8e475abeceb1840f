#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_audio.py tests/test_gpu_overlap.py tests/test_gpu_obs.py tests/test_gpu_step_graphs.py -x -q > gpurun_out/tests_audio_spectral.log 2>&1; tail -5 gpurun_out/tests_audio_spectral.log | cut -c1-250
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares > gpurun_out/bench_spectral_env.log 2>&1; grep '{"metric' gpurun_out/bench_spectral_env.log | sed 's/^[^{]*//' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default', d['value'], d['rollout_env_steps_per_s'], d['update_samples_per_s'], 'e2e', d['e2e']['value'], d['e2e']['rollout_env_steps_per_s'], 'trainable', d['trainable']['env_steps_per_s'], d['trainable']['rollout_env_steps_per_s'])"
timeout 900 python bench.py --config interactive --steps 2 --warmup 3 --no-cpu --no-shares > gpurun_out/bench_spectral_env_interactive.log 2>&1; grep '{"metric' gpurun_out/bench_spectral_env_interactive.log | sed 's/^[^{]*//' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('interactive', d['value'], d['rollout_env_steps_per_s'], d['update_samples_per_s'], 'e2e', d['e2e']['value'])"

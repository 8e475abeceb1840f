#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/wgrad_conv_bench.py 2>&1 | grep "gn_bwd\|wgrad" | cut -c1-250 | tee gpurun_out/wgrad_conv_bench_gg.txt
timeout 900 python -m pytest tests/test_gpu_nn.py tests/test_gpu_tc.py tests/test_gpu_avnav.py -x -q 2>&1 | tail -2
timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares --no-e2e > gpurun_out/bench_gg.log 2>&1; grep '{"metric' gpurun_out/bench_gg.log | sed 's/^[^{]*//' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('default', d['value'], d['rollout_env_steps_per_s'], d['update_samples_per_s'], 'trainable', d['trainable']['env_steps_per_s'], d['trainable']['rollout_env_steps_per_s'], d['trainable']['update_samples_per_s'])"

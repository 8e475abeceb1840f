#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_tc.py tests/test_gpu_nn.py tests/test_gpu_golden.py tests/test_gpu_avnav.py tests/test_gpu_policy.py tests/test_gpu_dialog.py -x -q 2>&1 | tail -2
for w in 0 1; do AVL_WIDE_STORES=$w timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares > gpurun_out/bench_wide$w.log 2>&1; grep '{"metric' gpurun_out/bench_wide$w.log | sed 's/^[^{]*//' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('wide $w default', d['value'], d['rollout_env_steps_per_s'], d['update_samples_per_s'], 'e2e', d['e2e']['value'], 'trainable', d['trainable']['env_steps_per_s'], d['trainable']['rollout_env_steps_per_s'], d['trainable']['update_samples_per_s'])"; done
AVL_WIDE_STORES=0 timeout 300 python tools/tma_conv_bench.py 2>&1 | tail -12 | cut -c1-60,100-200
AVL_WIDE_STORES=1 timeout 300 python tools/tma_conv_bench.py 2>&1 | tail -12 | cut -c1-60,100-200

#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_tc.py tests/test_gpu_nn.py tests/test_gpu_step_graphs.py tests/test_gpu_golden.py tests/test_gpu_policy.py -m gpu -x -q > gpurun_out/r02_tests_z.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_z.log
tail -6 gpurun_out/r02_tests_z.log | cut -c1-250
for pdl in 0 1; do AVL_PDL=$pdl timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu --no-eager --no-shares --regime frozen > gpurun_out/r02_bench_z_pdl$pdl.log 2>&1; grep '{"metric' gpurun_out/r02_bench_z_pdl$pdl.log | sed 's/^[^{]*//' | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('pdl=$pdl', d['value'], d['rollout_env_steps_per_s'], d['update_samples_per_s'], 'e2e', d['e2e']['value'], d['e2e']['rollout_env_steps_per_s'])"; grep -i "graphs disabled" gpurun_out/r02_bench_z_pdl$pdl.log | head -2; done

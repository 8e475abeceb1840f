timeout 300 python -m pytest tests/test_gpu_tc.py -q -k "3xtf32" > gpurun_out/pytest_x3.log 2>&1; echo "x3 pytest rc=$?"; tail -4 gpurun_out/pytest_x3.log | cut -c1-200
timeout 300 python tools/x3_gemm_bench.py 2>&1 | grep "3xTF32"
timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu29.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu29.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v19.log 2>&1; tail -1 gpurun_out/bench_r01_v19.log | cut -c1-300; grep -o '"rollout_env_steps_per_s": [0-9.]*, "update_samples_per_s": [0-9.]*, "e2e": {"value": [0-9.]*' gpurun_out/bench_r01_v19.log

timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu28.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu28.log | cut -c1-300
python tools/host_time.py > gpurun_out/host_time.log 2>&1; tail -4 gpurun_out/host_time.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v18.log 2>&1; tail -1 gpurun_out/bench_r01_v18.log | cut -c1-300; grep -o '"rollout_env_steps_per_s": [0-9.]*, "update_samples_per_s": [0-9.]*, "e2e": {"value": [0-9.]*' gpurun_out/bench_r01_v18.log;  grep -o '"gpu_launches": [0-9]*' gpurun_out/bench_r01_v18.log

timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu22.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu22.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v13.log 2>&1; tail -1 gpurun_out/bench_r01_v13.log | cut -c1-900
python tools/profile_step.py 150 1 > gpurun_out/profile_step_v13.log 2>&1; grep -v "Warn\|self.encoder\|_warn_once" gpurun_out/profile_step_v13.log | head -30
python tools/host_time.py > gpurun_out/host_time.log 2>&1; tail -2 gpurun_out/host_time.log

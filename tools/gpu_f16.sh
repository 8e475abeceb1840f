timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu26.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu26.log | cut -c1-300
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v16.log 2>&1; tail -1 gpurun_out/bench_r01_v16.log | cut -c1-1800
python tools/profile_step.py 150 1 > gpurun_out/profile_step_v16.log 2>&1; grep -v "Warn\|self.encoder\|_warn_once" gpurun_out/profile_step_v16.log | sed -n 1,45p

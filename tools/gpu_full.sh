timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu17.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu17.log
python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v8.log 2>&1; tail -1 gpurun_out/bench_r01_v8.log
python tools/profile_step.py 150 1 > gpurun_out/profile_step_v8.log 2>&1; grep -v "Warn\|self.encoder\|_warn_once" gpurun_out/profile_step_v8.log | head -48

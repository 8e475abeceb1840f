timeout 300 python tools/wgrad_probe.py > gpurun_out/wgrad_probe.log 2>&1; echo "probe rc=$?"; cat gpurun_out/wgrad_probe.log | tail -30
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu19.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu19.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v10.log 2>&1; tail -1 gpurun_out/bench_r01_v10.log
python tools/profile_step.py 150 1 > gpurun_out/profile_step_v10.log 2>&1; grep -v "Warn\|self.encoder\|_warn_once" gpurun_out/profile_step_v10.log | head -60

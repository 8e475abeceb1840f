timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu25.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu25.log | cut -c1-300
python tools/halo_f16_bench.py 4800 2>&1 | grep GN
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v15.log 2>&1; tail -1 gpurun_out/bench_r01_v15.log | cut -c1-1500
python tools/profile_step.py 150 1 > gpurun_out/profile_step_v15.log 2>&1; grep -v "Warn\|self.encoder\|_warn_once" gpurun_out/profile_step_v15.log | sed -n 22,45p

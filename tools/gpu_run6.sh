python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu7.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu7.log
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01_halo.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_r01_halo.log
python tools/profile_step.py 150 1 > gpurun_out/profile_step_halo.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/launches_r01_halo.csv python bench.py --rollout-steps 6 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch_halo.log 2>&1; echo "ncu rc=$?"

python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu4.log 2>&1; echo "pytest rc=$?"
python bench.py --steps 2 --warmup 3 > gpurun_out/bench_r01_tc1.log 2>&1; echo "bench rc=$?"
python bench.py --steps 2 --warmup 3 --tc-level 2 --no-cpu --no-e2e > gpurun_out/bench_r01_tc2.log 2>&1; echo "bench2 rc=$?"
python tools/profile_step.py 150 1 > gpurun_out/profile_step_tc1.log 2>&1
python tools/profile_step.py 150 2 > gpurun_out/profile_step_tc2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 30000 --csv --log-file gpurun_out/launches_r01_tc.csv python bench.py --rollout-steps 30 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/ncu_launch_tc.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/pytest_gpu4.log; tail -1 gpurun_out/bench_r01_tc1.log; tail -1 gpurun_out/bench_r01_tc2.log

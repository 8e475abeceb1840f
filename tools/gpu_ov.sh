timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu21.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu21.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v12.log 2>&1; tail -1 gpurun_out/bench_r01_v12.log
python tools/host_time.py > gpurun_out/host_time.log 2>&1; tail -3 gpurun_out/host_time.log

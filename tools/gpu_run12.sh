timeout 600 python -m pytest tests/test_gpu_nn.py tests/test_gpu_policy.py tests/test_gpu_tc.py -q -x > gpurun_out/pytest_gpu10.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu10.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke4.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke4.log
python tools/host_time.py 2>&1 | grep "steps:"
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r01_fused.log 2>&1; tail -1 gpurun_out/bench_r01_fused.log

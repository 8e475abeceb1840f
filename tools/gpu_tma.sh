timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu16.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu16.log
python tools/tc_gemm_bench.py 2>&1 | grep "tma  "
python tools/tc_conv_bench.py 4800 2>&1 | grep "tc.ca"
python tools/bench_interactive.py 32 30 3 2>&1 | grep "bench"
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r01_vec.log 2>&1; tail -1 gpurun_out/bench_r01_vec.log | cut -c1-900

timeout 300 python -m pytest tests/test_gpu_tc.py -q -k 3xtf32 > gpurun_out/pytest_x3.log 2>&1; echo "x3 pytest rc=$?"; tail -15 gpurun_out/pytest_x3.log
timeout 300 python tools/x3_gemm_bench.py > gpurun_out/x3_bench.log 2>&1; echo "x3 bench rc=$?"; cat gpurun_out/x3_bench.log | tail -40
timeout 600 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu18.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu18.log
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v9.log 2>&1; tail -1 gpurun_out/bench_r01_v9.log

timeout 600 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu27.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu27.log | cut -c1-300
python tools/halo_f16_bench.py 4800 2>&1 | grep -v "GN\|Trace\|File\|main\|f()\|assert\|\^\|Assertion"
timeout 600 python bench.py --steps 3 --warmup 3 > gpurun_out/bench_r01_v17.log 2>&1; tail -1 gpurun_out/bench_r01_v17.log | cut -c1-300; grep -o '"roofline": {[^}]*}' gpurun_out/bench_r01_v17.log | cut -c1-420

timeout 300 python -m pytest tests/test_gpu_tc.py -q -x > gpurun_out/pytest_gpu6.log 2>&1; echo "pytest rc=$?"
tail -5 gpurun_out/pytest_gpu6.log
timeout 120 python tools/tc_conv_bench.py 4800 > gpurun_out/tc_conv_bench4.log 2>&1
timeout 120 python tools/tc_conv_bench.py 64 >> gpurun_out/tc_conv_bench4.log 2>&1
HALO_ROWS=16 timeout 120 python tools/tc_conv_bench.py 4800 layer1 >> gpurun_out/tc_conv_bench4.log 2>&1
HALO_ROWS=4 timeout 120 python tools/tc_conv_bench.py 4800 layer1 >> gpurun_out/tc_conv_bench4.log 2>&1
grep halo gpurun_out/tc_conv_bench4.log

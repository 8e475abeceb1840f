python -m pytest tests/test_gpu_tc.py tests/test_gpu_dialog.py -q -x > gpurun_out/pytest_gpu5.log 2>&1; echo "pytest rc=$?"
tail -30 gpurun_out/pytest_gpu5.log
python tools/tc_conv_bench.py 4800 > gpurun_out/tc_conv_bench3.log 2>&1
python tools/tc_conv_bench.py 64 >> gpurun_out/tc_conv_bench3.log 2>&1
HALO_ROWS=16 python tools/tc_conv_bench.py 4800 layer1 >> gpurun_out/tc_conv_bench3.log 2>&1
HALO_ROWS=4 python tools/tc_conv_bench.py 4800 layer1 >> gpurun_out/tc_conv_bench3.log 2>&1
grep -v simt gpurun_out/tc_conv_bench3.log

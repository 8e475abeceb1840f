timeout 300 python -m pytest tests/test_gpu_tc.py tests/test_gpu_nn.py -q > gpurun_out/pytest_l.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_l.log | cut -c1-200
python tools/halo_f16_bench.py 4800 2>&1 | grep -v GN
python tools/tc_conv_bench.py 4800 "" 3 2>&1 | grep "tc.halo\|tc.ca"

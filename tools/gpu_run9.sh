timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_nn.py tests/test_gpu_policy.py -q -x > gpurun_out/pytest_gpu9.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/pytest_gpu9.log
python tools/tc_conv_bench.py 64 > gpurun_out/tc_conv_bench5.log 2>&1; grep "tc.ca\|tc.halo" gpurun_out/tc_conv_bench5.log
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r01_splitk.log 2>&1; tail -1 gpurun_out/bench_r01_splitk.log
python tools/profile_step.py 150 1 > gpurun_out/profile_step_splitk.log 2>&1

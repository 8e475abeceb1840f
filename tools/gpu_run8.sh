timeout 300 python -m pytest tests/test_gpu_nn.py -q -x > gpurun_out/pytest_gpu8.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu8.log
ncu --set full --import-source on --clock-control none -k regex:tc_gemm_kernel -s 3 -c 1 -o gpurun_out/prof_tcgemm_ws_layer4_r01 -f python tools/tc_conv_bench.py 4800 layer4 2 > gpurun_out/ncu_ws.log 2>&1; echo "ncu rc=$?"
python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/bench_r01_gnc.log 2>&1; tail -1 gpurun_out/bench_r01_gnc.log

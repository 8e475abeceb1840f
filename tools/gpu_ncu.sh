python tools/gn_bench.py 4800 > gpurun_out/gn_bench.log 2>&1; python tools/gn_bench.py 64 >> gpurun_out/gn_bench.log 2>&1; cat gpurun_out/gn_bench.log | grep "B="
ncu --set full --import-source on --clock-control none -k regex:tc_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo3_layer1_r01 -f python tools/tc_conv_bench.py 4800 layer1 2 > gpurun_out/ncu_a.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none -k regex:gn_cluster -s 2 -c 1 -o gpurun_out/prof_gncluster_r01 -f python tools/gn_bench.py 4800 > gpurun_out/ncu_b.log 2>&1; echo "ncu rc=$?"
ncu --set full --clock-control none -k regex:tc_gemm_tma -s 70 -c 1 -o gpurun_out/prof_tma_gemm_r01 -f python tools/tc_gemm_bench.py > gpurun_out/ncu_c.log 2>&1; echo "ncu rc=$?"

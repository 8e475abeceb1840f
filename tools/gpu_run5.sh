ncu --set full --import-source on --clock-control none -k regex:tc_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo2_layer1_r01 -f python tools/tc_conv_bench.py 4800 layer1 2 > gpurun_out/ncu_halo2.log 2>&1; echo "ncu rc=$?"
ncu --set full --import-source on --clock-control none -k regex:tc_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo2_layer2_r01 -f python tools/tc_conv_bench.py 4800 32-\>32 2 >> gpurun_out/ncu_halo2.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_halo2.log

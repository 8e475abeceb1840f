ncu --set full --import-source on --clock-control none -k regex:tc_conv_halo -s 3 -c 1 -o gpurun_out/prof_halo_layer1_r01 -f python tools/tc_conv_bench.py 4800 layer1 2 > gpurun_out/ncu_halo.log 2>&1; echo "ncu rc=$?"
tail -5 gpurun_out/ncu_halo.log

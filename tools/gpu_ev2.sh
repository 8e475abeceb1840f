#!/bin/bash
mkdir -p gpurun_out
ncu --set full --import-source on --clock-control none -k regex:tc_gemm_kernel -s 6 -c 1 -o gpurun_out/prof_splitk_cluster_layer4_b64_r01 -f python tools/tc_conv_bench.py 64 "layer4" 3 > gpurun_out/ncu_m.log 2>&1; echo "ncu rc=$?"

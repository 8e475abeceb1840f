#!/bin/bash
# ncu --set full of the kernels added in the second half of round 2
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "tma_im2col or tc_conv or splitk" > gpurun_out/r02_tests_w.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_w.log
tail -4 gpurun_out/r02_tests_w.log | cut -c1-250
timeout 600 python tools/tma_conv_bench.py 4800 > gpurun_out/r02_tma_conv_bench_w.txt 2>&1; grep "^B=" gpurun_out/r02_tma_conv_bench_w.txt
for p in tma_conv_l4 tma_conv_l4_b64 attn_tc_fwd; do
  python tools/ncu_probe.py $p > gpurun_out/ncu_plain_$p.log 2>&1 || { echo "plain run of $p failed"; tail -5 gpurun_out/ncu_plain_$p.log; }
done
ncu --set full --clock-control none --import-source on -k regex:tc_gemm_tma_kernel -s 2 -c 1 -o gpurun_out/r02_tma_conv_l4 python tools/ncu_probe.py tma_conv_l4 > gpurun_out/ncu_tma_conv_l4.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_tma_splitk_kernel -s 2 -c 1 -o gpurun_out/r02_tma_splitk_l4_b64 python tools/ncu_probe.py tma_conv_l4_b64 > gpurun_out/ncu_tma_splitk.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_self_fwd_tc_kernel -s 2 -c 1 -o gpurun_out/r02_attn_tc_fwd python tools/ncu_probe.py attn_tc_fwd > gpurun_out/ncu_attn_fwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:attn_self_bwd_tc_kernel -s 2 -c 1 -o gpurun_out/r02_attn_tc_bwd python tools/ncu_probe.py attn_tc_bwd > gpurun_out/ncu_attn_bwd.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -6
tail -3 gpurun_out/ncu_tma_conv_l4.log

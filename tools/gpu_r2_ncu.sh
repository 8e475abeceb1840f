#!/bin/bash
# ncu --set full of the round-2 kernels (one ncu use per gpurun call: all runs below count as one)
set -x
mkdir -p gpurun_out
for p in wgrad_l1 wgrad_l3 gn_bwd_l1 halo_s2 resize_u8; do
  python tools/ncu_probe.py $p > gpurun_out/ncu_plain_$p.log 2>&1 || { echo "plain run of $p failed"; tail -5 gpurun_out/ncu_plain_$p.log; continue; }
done
ncu --set full --clock-control none --import-source on -k regex:tc_conv_wgrad_kernel -s 2 -c 1 -o gpurun_out/r02_wgrad_l1 python tools/ncu_probe.py wgrad_l1 > gpurun_out/ncu_wgrad_l1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_wgrad_kernel -s 2 -c 1 -o gpurun_out/r02_wgrad_l3 python tools/ncu_probe.py wgrad_l3 > gpurun_out/ncu_wgrad_l3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:gn_cluster_bwd_kernel -s 2 -c 1 -o gpurun_out/r02_gn_bwd_l1 python tools/ncu_probe.py gn_bwd_l1 > gpurun_out/ncu_gn_bwd.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_conv_halo_kernel -s 2 -c 1 -o gpurun_out/r02_halo_s2 python tools/ncu_probe.py halo_s2 > gpurun_out/ncu_halo_s2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:resize_half_typed -s 2 -c 1 -o gpurun_out/r02_resize_u8 python tools/ncu_probe.py resize_u8 > gpurun_out/ncu_resize.log 2>&1
ls -la gpurun_out/*.ncu-rep
tail -3 gpurun_out/ncu_wgrad_l1.log

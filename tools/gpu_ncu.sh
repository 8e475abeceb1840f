#!/bin/bash
# ncu --set full of one kernel through tools/ncu_probe.py (one capture per gpurun call; the plain run comes first):
#   gpurun --timeout 1200 -- 'bash tools/gpu_ncu.sh tma_conv_l4 tc_gemm_tma_kernel'
# then here:  python tools/ncu_summary.py gpurun_out/<probe>.ncu-rep > profiles/<name>_ncu_full.txt
probe=$1; regex=$2
set -x
mkdir -p gpurun_out
python tools/ncu_probe.py $probe > gpurun_out/ncu_plain_$probe.log 2>&1 || { echo "plain run of $probe failed"; tail -5 gpurun_out/ncu_plain_$probe.log; exit 1; }
ncu --set full --clock-control none --import-source on -k regex:$regex -s 2 -c 1 -o gpurun_out/$probe python tools/ncu_probe.py $probe > gpurun_out/ncu_$probe.log 2>&1
ls -la gpurun_out/$probe.ncu-rep; tail -3 gpurun_out/ncu_$probe.log

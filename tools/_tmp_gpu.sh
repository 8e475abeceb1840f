#!/bin/bash
mkdir -p gpurun_out
cp avlen_b200/libavlen_b200.so /tmp/lib_after.so
echo "== after (batched fp16 loads)"; timeout 300 python tools/halo_f16_bench.py 4800 2>&1 | grep "GN " | tee gpurun_out/gn_f16_batched.txt
cp _ab/lib_before.so avlen_b200/libavlen_b200.so
echo "== before"; timeout 300 python tools/halo_f16_bench.py 4800 2>&1 | grep "GN " | tee gpurun_out/gn_f16_before.txt
cp /tmp/lib_after.so avlen_b200/libavlen_b200.so

#!/bin/bash
mkdir -p gpurun_out
for st in 0 3 4; do
  TC_STAGES=$st timeout 300 python tools/tc_conv_bench.py 4800 > gpurun_out/tc_conv_b4800_st$st.log 2>&1
  echo "== stages $st"; grep -E "tc.ca" gpurun_out/tc_conv_b4800_st$st.log
done

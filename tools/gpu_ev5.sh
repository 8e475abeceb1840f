#!/bin/bash
mkdir -p gpurun_out
timeout 200 ncu --set full --import-source on --clock-control none -k regex:attn_self_fwd -s 140 -c 1 -o gpurun_out/prof_attn_self_fwd_rollout_r01 -f python tools/attn_rollout_probe.py > gpurun_out/ncu_p.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_p.log

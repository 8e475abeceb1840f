#!/bin/bash
mkdir -p gpurun_out
AVL_TRACE_UPDATE=gpurun_out/trace_update_v23.json timeout 600 python tools/profile_step.py 150 1 > gpurun_out/profile_step_v23b.log 2>&1; echo "profile rc=$?"
gzip -f gpurun_out/trace_update_v23.json; ls -la gpurun_out/trace_update_v23.json.gz

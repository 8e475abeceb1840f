#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python tools/e2e_step_probe.py > gpurun_out/r02_e2e_step_probe_r.txt 2>&1; grep -v Warn gpurun_out/r02_e2e_step_probe_r.txt | grep -v "^ " | tail -30; grep -c "^ " gpurun_out/r02_e2e_step_probe_r.txt

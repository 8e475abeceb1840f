#!/bin/bash
set -x
mkdir -p gpurun_out
AVL_POLICY=interactive timeout 900 python tools/profile_step.py 150 > gpurun_out/r02_profile_interactive_cc.txt 2>&1; grep -A32 "10 rollout steps" gpurun_out/r02_profile_interactive_cc.txt | cut -c1-170

#!/bin/bash
# evidence for the cluster split-K conv, the channel-split audio kernel, config-3/4 benches
mkdir -p gpurun_out
timeout 300 python tools/tc_conv_bench.py 64 > gpurun_out/tc_conv_bench_b64_v23.log 2>&1; grep "B=64" gpurun_out/tc_conv_bench_b64_v23.log | grep "tc.halo"
timeout 300 python tools/audio_sweep.py --envs 32 64 128 256 1024 4096 > gpurun_out/audio_sweep_v2.jsonl 2>gpurun_out/audio_sweep_v2.err; cat gpurun_out/audio_sweep_v2.jsonl | cut -c1-200
timeout 300 python tools/audio_sweep.py --envs 64 1024 --audiogoal 0 >> gpurun_out/audio_sweep_v2.jsonl 2>>gpurun_out/audio_sweep_v2.err
timeout 300 python tools/audio_sweep.py --envs 64 1024 --distractor 1 >> gpurun_out/audio_sweep_v2.jsonl 2>>gpurun_out/audio_sweep_v2.err
timeout 300 python tools/audio_sweep.py --envs 64 1024 --lens 4000 8000 >> gpurun_out/audio_sweep_v2.jsonl 2>>gpurun_out/audio_sweep_v2.err
timeout 600 python tools/bench_interactive.py > gpurun_out/bench_interactive_v23.log 2>&1; tail -4 gpurun_out/bench_interactive_v23.log
ncu --set full --import-source on --clock-control none -k regex:tc_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_splitk_cluster_layer4_b64_r01 -f python tools/tc_conv_bench.py 64 "layer4" 3 > gpurun_out/ncu_m.log 2>&1; echo "ncu rc=$?"
ncu --set full --import-source on --clock-control none -k regex:audio_render -s 2 -c 1 -o gpurun_out/prof_audio_split64_r01 -f python tools/audio_sweep.py --envs 64 --iters 3 > gpurun_out/ncu_n.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -3

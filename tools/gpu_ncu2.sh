python tools/attn_bench.py > gpurun_out/attn_bench.log 2>&1; cat gpurun_out/attn_bench.log | grep -v Warn
ncu --set full --import-source on --clock-control none -k regex:tc_gemm_3x -s 17 -c 1 -o gpurun_out/prof_gemm3x_r01 -f python tools/x3_gemm_bench.py > gpurun_out/ncu_d.log 2>&1; echo "ncu rc=$?"
ncu --set full --import-source on --clock-control none -k regex:tc_wgrad_3x -s 3 -c 1 -o gpurun_out/prof_wgrad3x_r01 -f python tools/wgrad_probe.py > gpurun_out/ncu_e.log 2>&1; echo "ncu rc=$?"
ncu --set full --import-source on --clock-control none -k regex:attn_self_bwd -s 2 -c 1 -o gpurun_out/prof_attn_bwd_r01 -f python tools/attn_bench.py > gpurun_out/ncu_f.log 2>&1; echo "ncu rc=$?"
ncu --set full --import-source on --clock-control none -k regex:attn_self_fwd -s 3 -c 1 -o gpurun_out/prof_attn_fwd_r01 -f python tools/attn_bench.py > gpurun_out/ncu_g.log 2>&1; echo "ncu rc=$?"
ls -la gpurun_out/*.ncu-rep | tail -5

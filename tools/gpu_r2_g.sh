#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_tc.py tests/test_gpu_nn.py tests/test_gpu_tc_bwd.py -m gpu -x -q > gpurun_out/r02_tests_g.log 2>&1; echo "tests rc=$?" >> gpurun_out/r02_tests_g.log
tail -5 gpurun_out/r02_tests_g.log | cut -c1-200
python - <<'PY' > gpurun_out/r02_halo_stride2_bench.txt 2>&1
import torch, sys
sys.path.insert(0, '.')
from avlen_b200 import nn as K, _lib
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def t(fn, it=5):
    for _ in range(2): fn()
    ts=[]
    for _ in range(it):
        flush.zero_(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts)//2]
for B in (4800, 64):
    for name,H,C,Co in (("l2e 16->32 @64",64,16,32),("l3e 32->64 @32",32,32,64),("l4e 64->128 @16",16,64,128)):
        x=torch.randn(B,H,H,C,device="cuda"); w=torch.randn(Co,C,3,3,device="cuda")/(C*9)**0.5
        outs=[]
        for on in (0,1):
            _lib.lib().avl_set_tc_conv_halo_stride2(on)
            ms=t(lambda: K._conv2d_raw(x,w,None,2,1))
            outs.append(K._conv2d_raw(x,w,None,2,1).clone())
            print(f"B={B} {name} halo_stride2={on}: {ms:.3f} ms", flush=True)
        _lib.lib().avl_set_tc_conv_halo_stride2(1)
        print("   max rel diff halo-s2 vs im2col:", float((outs[0]-outs[1]).abs().max()/outs[0].abs().max()), flush=True)
PY
cat gpurun_out/r02_halo_stride2_bench.txt
timeout 600 python bench.py --steps 3 --warmup 3 --regime frozen --no-cpu --no-eager --no-shares > gpurun_out/r02_bench_g_frozen.log 2>&1; tail -1 gpurun_out/r02_bench_g_frozen.log | cut -c1-700
timeout 600 python bench.py --steps 3 --warmup 3 --regime trainable --no-cpu --no-eager --no-shares --no-e2e > gpurun_out/r02_bench_g_trainable.log 2>&1; tail -1 gpurun_out/r02_bench_g_trainable.log | cut -c1-700

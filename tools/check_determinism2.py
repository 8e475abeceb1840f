import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avlen_b200 import nn as K

g = torch.Generator().manual_seed(1)
def rnd(*s): return torch.randn(*s, generator=g).cuda()
def same(f, name):
    a = f().clone(); b = f().clone(); c = f().clone()
    torch.cuda.synchronize()
    d = max(float((a - b).abs().max()), float((a - c).abs().max()))
    print(f"{name:50s} maxdiff {d:.3e}  (max {float(a.abs().max()):.3f})", flush=True)

for n in (2, 64):
    x16 = rnd(n, 64, 64, 16); w16 = rnd(16, 16, 3, 3) / 12
    x4 = rnd(n, 64, 64, 4); w4 = rnd(16, 4, 7, 7) / 14
    x64 = rnd(n, 16, 16, 64); w64 = rnd(64, 64, 3, 3) / 24
    x128 = rnd(n, 8, 8, 128); w128 = rnd(128, 128, 3, 3) / 34; wfc = rnd(64, 128, 8, 8) / 90
    x32 = rnd(n, 32, 32, 16); w32 = rnd(32, 16, 3, 3) / 12; wd = rnd(32, 16, 1, 1) / 4
    ga, be = rnd(16).abs() + 0.5, rnd(16)
    K.set_tensor_cores(1)
    for sk in (1, 0):
        K._lib.lib().avl_set_tc_splitk(sk)
        same(lambda: K.conv2d(x16, w16, None, 1, 1), f"n={n} sk={sk} halo conv 3x3 16->16")
        same(lambda: K.conv2d(x4, w4, None, 1, 3), f"n={n} sk={sk} halo conv 7x7 4->16")
        same(lambda: K.conv2d(x64, w64, None, 1, 1), f"n={n} sk={sk} generic conv 64->64 @16")
        same(lambda: K.conv2d(x128, w128, None, 1, 1), f"n={n} sk={sk} generic conv 128->128 @8")
        same(lambda: K.conv2d(x128, wfc, None, 1, 0), f"n={n} sk={sk} fc 8x8x128->64")
        same(lambda: K.conv2d(x32, w32, None, 2, 1), f"n={n} sk={sk} generic conv s2 16->32")
        same(lambda: K.conv2d(x32, wd, None, 2, 0), f"n={n} sk={sk} generic conv 1x1 s2 16->32")
    same(lambda: K.groupnorm(x16, ga, be, 16, 1e-5, relu=True), f"n={n} gn cluster 64x64x16")
    same(lambda: K.groupnorm(x16.clone(), ga, be, 16, 1e-5, relu=True, residual=x16), f"n={n} gn cluster + residual")
    K.set_tensor_cores(0)
    same(lambda: K.conv2d(x64, w64, None, 1, 1), f"n={n} SIMT conv 64->64 @16")
    K.set_tensor_cores(1)

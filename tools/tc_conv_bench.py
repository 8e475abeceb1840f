"""Times the encoder convolution shapes of custom_resnet18 at a given batch on the tensor-core and SIMT paths
(CUDA events, L2 flushed between iterations) and prints achieved TFLOP/s and algorithmic GB/s per layer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200 import nn as K

SHAPES = [  # H, W, C, Cout, KH, KW, stride, pad
    ("conv1 7x7 4->16 @64", 64, 64, 4, 16, 7, 7, 1, 3),
    ("layer1 3x3 16->16 @64", 64, 64, 16, 16, 3, 3, 1, 1),
    ("layer2 3x3 16->32 s2", 64, 64, 16, 32, 3, 3, 2, 1),
    ("layer2 3x3 32->32 @32", 32, 32, 32, 32, 3, 3, 1, 1),
    ("layer3 3x3 32->64 s2", 32, 32, 32, 64, 3, 3, 2, 1),
    ("layer4 3x3 64->128 s2", 16, 16, 64, 128, 3, 3, 2, 1),
    ("layer3 3x3 64->64 @16", 16, 16, 64, 64, 3, 3, 1, 1),
    ("layer4 3x3 128->128 @8", 8, 8, 128, 128, 3, 3, 1, 1),
    ("fc 8x8x128->64", 8, 8, 128, 64, 8, 8, 1, 0),
]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4800
    only = sys.argv[2] if len(sys.argv) > 2 else None
    iters = int(sys.argv[3]) if len(sys.argv) > 3 else 5
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    if os.environ.get("TC_STAGES"):
        K._lib.lib().avl_set_tc_stages(int(os.environ["TC_STAGES"]))
    for name, H, W, C, Co, KH, KW, s, p in SHAPES:
        if only and only not in name:
            continue
        x = torch.randn(B, H, W, C, device="cuda")
        w = torch.randn(Co, C, KH, KW, device="cuda") / (C * KH * KW) ** 0.5
        OH, OW = K.conv_out(H, KH, s, p), K.conv_out(W, KW, s, p)
        flops = 2.0 * B * OH * OW * Co * C * KH * KW
        byts = 4.0 * (x.numel() + B * OH * OW * Co)
        for level in (4, 3, 2, 1, 0):  # 4: im2col into SWIZZLE_NONE planes; 3: halo-strip kernel; 2: im2col cp.async.ca; 1: cp.async.cg; 0: fp32 SIMT
            K.set_tensor_cores(min(level, 1))
            K._lib.lib().avl_set_tc_conv_l1(1 if level >= 2 else 0)
            K.set_conv_halo(1 if level == 3 else 0, int(os.environ.get("HALO_ROWS", "0")))
            K._lib.lib().avl_set_tc_swizzle(0 if level == 4 else 1)
            for _ in range(2):
                K.conv2d(x, w, None, s, p)
            ts = []
            for _ in range(iters):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                K.conv2d(x, w, None, s, p)
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            print(f"B={B} {name:28s} {('simt', 'tc.cg', 'tc.ca', 'tc.halo', 'tc.nosw')[level]} {ms:8.3f} ms  {flops / ms / 1e9:8.2f} TFLOP/s  "
                  f"{byts / ms / 1e6:8.1f} GB/s", flush=True)
    K.set_tensor_cores(1)
    K._lib.lib().avl_set_tc_conv_l1(1)
    K.set_conv_halo(1, 0)
    K._lib.lib().avl_set_tc_swizzle(1)


if __name__ == "__main__":
    main()

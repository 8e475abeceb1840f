"""Times the deep encoder convolutions at update batch on the TMA im2col kernel and on the cp.async im2col kernel
(CUDA events, L2 flushed between iterations, median of 7): achieved TFLOP/s (TF32) and algorithmic GB/s per layer."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200 import nn as K

SHAPES = [  # name, H, W, C, Cout, k, stride, pad
    ("layer2 1x1 s2 16->32 @64", 64, 64, 16, 32, 1, 2, 0),
    ("layer3 3x3 s2 32->64 @32", 32, 32, 32, 64, 3, 2, 1),
    ("layer3 1x1 s2 32->64 @32", 32, 32, 32, 64, 1, 2, 0),
    ("layer3 3x3 64->64 @16", 16, 16, 64, 64, 3, 1, 1),
    ("layer4 3x3 s2 64->128 @16", 16, 16, 64, 128, 3, 2, 1),
    ("layer4 1x1 s2 64->128 @16", 16, 16, 64, 128, 1, 2, 0),
    ("layer4 3x3 128->128 @8", 8, 8, 128, 128, 3, 1, 1),
    ("belief l1 3x3 64->64 @17x7", 17, 7, 64, 64, 3, 1, 1),
    ("belief l2 3x3 128->128 @9x4", 9, 4, 128, 128, 3, 1, 1),
    ("belief l3 3x3 256->256 @5x2", 5, 2, 256, 256, 3, 1, 1),
    ("belief l4 3x3 512->512 @3x1", 3, 1, 512, 512, 3, 1, 1),
]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4800
    lib = K._lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    K.set_tensor_cores(1)
    for name, H, W, C, Co, k, s, p in SHAPES:
        x = torch.randn(B, H, W, C, device="cuda")
        w = torch.randn(Co, C, k, k, device="cuda") / (C * k * k) ** 0.5
        OH, OW = K.conv_out(H, k, s, p), K.conv_out(W, k, s, p)
        flops = 2.0 * B * OH * OW * Co * C * k * k
        byts = 4.0 * (x.numel() + B * OH * OW * Co + w.numel())
        out = torch.empty(B, OH, OW, Co, device="cuda")
        row = [f"B={B} {name:30s}"]
        for tma in (0, 1):
            lib.avl_set_tc_conv_tma(tma)
            n0 = int(lib.avl_tc_conv_tma_count())
            for _ in range(2):
                K.conv2d(x, w, None, s, p, out=out.view(-1, Co))
            took = int(lib.avl_tc_conv_tma_count()) - n0
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                K.conv2d(x, w, None, s, p, out=out.view(-1, Co))
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[3]
            row.append(f"{'tma' if took else 'cp.async'} {ms:7.3f} ms {flops / ms / 1e9:7.1f} TFLOP/s {byts / ms / 1e6:7.1f} GB/s")
        print("  |  ".join(row), flush=True)
    lib.avl_set_tc_conv_tma(1)


if __name__ == "__main__":
    main()

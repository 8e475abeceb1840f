"""Diagnostic: steady-state duration of the rollout-batch convolutions that run on the generic tensor-core kernel
(belief-predictor resnet18 on the 65x26x2 spectrogram, deep layers of custom_resnet18) — 40 launches back to back on
one stream (each waits for the previous one, as in the per-network chains of a rollout step; no L2 flush: weights stay
resident as they do in a step), under the split-K / ring-depth switches of the C ABI."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200 import nn as K

SHAPES = [  # H, W, C, Cout, k, stride, pad
    ("belief stem 7x7 s2 4->64", 65, 26, 4, 64, 7, 2, 3),
    ("belief l1 3x3 64->64 @17x7", 17, 7, 64, 64, 3, 1, 1),
    ("belief l2 3x3 s2 64->128", 17, 7, 64, 128, 3, 2, 1),
    ("belief l2 1x1 s2 64->128", 17, 7, 64, 128, 1, 2, 0),
    ("belief l2 3x3 128->128 @9x4", 9, 4, 128, 128, 3, 1, 1),
    ("belief l3 3x3 s2 128->256", 9, 4, 128, 256, 3, 2, 1),
    ("belief l3 3x3 256->256 @5x2", 5, 2, 256, 256, 3, 1, 1),
    ("belief l4 3x3 s2 256->512", 5, 2, 256, 512, 3, 2, 1),
    ("belief l4 3x3 512->512 @3x1", 3, 1, 512, 512, 3, 1, 1),
    ("visual l3 3x3 s2 32->64", 32, 32, 32, 64, 3, 2, 1),
    ("visual l4 3x3 s2 64->128", 16, 16, 64, 128, 3, 2, 1),
    ("visual l4 1x1 s2 64->128", 16, 16, 64, 128, 1, 2, 0),
    ("visual l4 3x3 128->128 @8", 8, 8, 128, 128, 3, 1, 1),
]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    lib = K._lib.lib()
    K.set_tensor_cores(1)
    for name, H, W, C, Co, k, s, p in SHAPES:
        x = torch.randn(B, H, W, C, device="cuda")
        w = torch.randn(Co, C, k, k, device="cuda") / (C * k * k) ** 0.5
        row = [f"B={B} {name:30s}"]
        for label, stages, splitk, cluster, tma in (("tma", 0, 1, 1, 1), ("tma ring4", 4, 1, 1, 1), ("cp.async", 0, 1, 1, 0),
                                                    ("cp.async ring4", 4, 1, 1, 0)):
            lib.avl_set_tc_conv_tma(tma)
            lib.avl_set_tc_stages(stages)
            lib.avl_set_tc_splitk(splitk)
            lib.avl_set_tc_splitk_cluster(cluster)
            out = torch.empty(B, K.conv_out(H, k, s, p), K.conv_out(W, k, s, p), Co, device="cuda")
            for _ in range(3):
                K.conv2d(x, w, None, s, p, out=out.view(-1, Co))
            torch.cuda.synchronize()
            graph = torch.cuda.CUDAGraph()   # device-side chain: the host's ~17 us per ctypes call must not be what is timed
            with torch.cuda.graph(graph):
                for _ in range(40):
                    K.conv2d(x, w, None, s, p, out=out.view(-1, Co))
            graph.replay()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            graph.replay()
            e1.record()
            torch.cuda.synchronize()
            row.append(f"{label} {e0.elapsed_time(e1) / 40 * 1e3:6.1f} us")
        print("  ".join(row), flush=True)
    lib.avl_set_tc_conv_tma(1)
    lib.avl_set_tc_stages(0)
    lib.avl_set_tc_splitk(1)
    lib.avl_set_tc_splitk_cluster(1)


if __name__ == "__main__":
    main()

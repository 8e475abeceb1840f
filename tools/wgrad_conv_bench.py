"""Per-layer timing of the convolution backward kernels at the update-minibatch batch (diagnostic):
tensor-core weight gradient (csrc/conv_bwd_tc.cu), data gradient through the forward kernels, GroupNorm backward."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200 import nn as K

LAYERS = [("stem7x7_4->16@64", 64, 4, 16, 7, 1, 3), ("l1_3x3_16->16@64", 64, 16, 16, 3, 1, 1),
          ("l2e_3x3s2_16->32@64", 64, 16, 32, 3, 2, 1), ("l2s_1x1s2_16->32@64", 64, 16, 32, 1, 2, 0),
          ("l2_3x3_32->32@32", 32, 32, 32, 3, 1, 1), ("l3e_3x3s2_32->64@32", 32, 32, 64, 3, 2, 1),
          ("l3_3x3_64->64@16", 16, 64, 64, 3, 1, 1), ("l4e_3x3s2_64->128@16", 16, 64, 128, 3, 2, 1),
          ("l4_3x3_128->128@8", 8, 128, 128, 3, 1, 1)]


def timeit(fn, flush, iters=5):
    for _ in range(2):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return sorted(ts)[len(ts) // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4800
    K.set_tensor_cores(1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, H, C, Co, k, s, p in LAYERS:
        x = torch.randn(B, H, H, C, device="cuda")
        w = torch.randn(Co, min(C, 3) if C == 4 else C, k, k, device="cuda") / (C * k * k) ** 0.5
        OH = (H + 2 * p - k) // s + 1
        gy = torch.randn(B, OH, OH, Co, device="cuda")
        ms_w = timeit(lambda: K.conv2d_wgrad_tc(x, gy, tuple(w.shape), s, p), flush)
        ms_d = None
        if C != 4:
            ms_d = timeit(lambda: K.conv2d_dgrad_tc(gy, w, H, H, s, p), flush)
        bytes_w = (x.numel() + gy.numel()) * 4
        flops = 2.0 * B * OH * OH * Co * C * k * k
        print(json.dumps({"layer": name, "batch": B, "wgrad_ms": round(ms_w, 3), "wgrad_GBps": round(bytes_w / ms_w / 1e6, 1),
                          "wgrad_TFLOPs": round(flops / ms_w / 1e9, 1), "dgrad_ms": None if ms_d is None else round(ms_d, 3),
                          "dgrad_GBps": None if ms_d is None else round(bytes_w / ms_d / 1e6, 1)}), flush=True)
        # GroupNorm backward on the layer's output
        xg = torch.randn(B, OH, OH, Co, device="cuda", requires_grad=True)
        gam = torch.ones(Co, device="cuda", requires_grad=True)
        bet = torch.zeros(Co, device="cuda", requires_grad=True)
        y = K.groupnorm(xg, gam, bet, 16, 1e-5, relu=True)

        def gnb():
            y.backward(gy, retain_graph=True)
        ms_g = timeit(gnb, flush)
        print(json.dumps({"layer": name, "gn_bwd_ms": round(ms_g, 3), "gn_bwd_GBps_4tensors": round(4 * gy.numel() * 4 / ms_g / 1e6, 1)}),
              flush=True)
        del x, gy, xg, y


if __name__ == "__main__":
    main()

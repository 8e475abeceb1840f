"""Times GroupNorm(16)+ReLU on the custom_resnet18 activation shapes: single-pass cluster kernel vs two-pass kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from avlen_b200 import nn as K

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4800
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for H, C in ((64, 16), (32, 32), (16, 64), (8, 128)):
        x = torch.randn(B, H, H, C, device="cuda")
        ga, be = torch.rand(C, device="cuda") + 0.5, torch.randn(C, device="cuda")
        for mode in (1, 0):
            K.set_groupnorm_cluster(mode)
            for _ in range(2):
                K.groupnorm(x, ga, be, 16, 1e-5, relu=True)
            ts = []
            for _ in range(7):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); K.groupnorm(x, ga, be, 16, 1e-5, relu=True); e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[3]
            print(f"B={B} {H}x{H}x{C} {'cluster ' if mode else 'two-pass'} {ms * 1e3:8.1f} us  {2 * x.numel() * 4 / ms / 1e6:8.1f} GB/s (read + write)", flush=True)
    K.set_groupnorm_cluster(1)

if __name__ == "__main__":
    main()

"""Times the dense tcgen05 GEMM (avl_tc_gemm) on the shapes of the CLIP text tower / SMT linears, TMA-fed kernel vs
cp.async kernel (CUDA events, median of 7, L2 flushed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200 import _lib
from avlen_b200 import nn as K

SHAPES = [(2541, 1536, 512), (2541, 512, 512), (2541, 2048, 512), (2541, 512, 2048), (1078, 1536, 512),
          (9600, 256, 288), (9600, 768, 256), (360000, 256, 256), (360000, 768, 256), (4800, 64, 8192)]


def main():
    lib = _lib.lib()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for M, N, Kd in SHAPES:
        x = torch.randn(M, Kd, device="cuda")
        w = torch.randn(N, Kd, device="cuda") / Kd ** 0.5
        b = torch.randn(N, device="cuda")
        out = torch.empty(M, N, device="cuda")
        ref = None
        for tma in (1, 0):
            lib.avl_set_tc_tma(tma)
            ts = []
            for _ in range(9):
                flush.zero_()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                _lib.call("avl_tc_gemm", x.data_ptr(), Kd, w.data_ptr(), out.data_ptr(), N, M, N, Kd, None, b.data_ptr(), None,
                          0, 0, None, _lib.stream())
                e1.record()
                torch.cuda.synchronize()
                ts.append(e0.elapsed_time(e1))
            ms = sorted(ts)[len(ts) // 2]
            if ref is None:
                ref = out.clone()
            err = float((out - ref).abs().max())
            print(f"M={M:7d} N={N:5d} K={Kd:5d} {'tma     ' if tma else 'cp.async'} {ms * 1e3:9.1f} us  "
                  f"{2.0 * M * N * Kd / ms / 1e9:8.2f} TFLOP/s  maxdiff vs tma {err:.2e}", flush=True)
    lib.avl_set_tc_tma(1)


if __name__ == "__main__":
    main()

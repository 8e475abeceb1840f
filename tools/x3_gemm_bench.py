"""Times the fp32-accurate 3xTF32 tcgen05 GEMM (avl_tc_gemm_3x) against the fp32 SIMT GEMM (avl_gemm) and cuBLAS fp32
(torch.matmul, allow_tf32 off) on the scene-memory transformer's linear shapes (CUDA events, median of 7, L2 flushed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200 import _lib
from avlen_b200 import nn as K  # noqa: F401

SHAPES = [(9600, 256, 276, 0), (9600, 768, 256, 0), (360000, 256, 256, 0), (360000, 768, 256, 0), (360000, 512, 256, 0),
          (360000, 256, 276, 0), (360000, 256, 768, 1), (360000, 256, 512, 1), (360000, 256, 256, 1), (5000, 256, 256, 0)]


def timeit(fn, flush):
    ts = []
    for _ in range(7):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def main():
    lib = _lib.lib()
    torch.backends.cuda.matmul.allow_tf32 = False
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for M, N, Kd, bt in SHAPES:
        x = torch.randn(M, Kd, device="cuda")
        w = torch.randn(N, Kd, device="cuda") / Kd ** 0.5
        b = torch.randn(N, device="cuda")
        wb = w.t().contiguous() if bt else w
        ldb = N if bt else Kd
        out = torch.empty(M, N, device="cuda")
        ref = torch.addmm(b.double(), x[:4096].double(), w.double().t()).float()

        def f3():
            rc = lib.avl_tc_gemm_3x(x.data_ptr(), Kd, wb.data_ptr(), ldb, bt, out.data_ptr(), N, M, N, Kd, b.data_ptr(), None,
                                    0, 0, None, _lib.stream())
            assert rc == 0, rc

        def fs():
            if bt:
                _lib.call("avl_gemm", x.data_ptr(), Kd, 1, wb.data_ptr(), 1, N, out.data_ptr(), N, M, N, Kd, b.data_ptr(), 0, 0, 1,
                          _lib.stream())
            else:
                _lib.call("avl_gemm", x.data_ptr(), Kd, 1, wb.data_ptr(), Kd, 1, out.data_ptr(), N, M, N, Kd, b.data_ptr(), 0, 0, 1,
                          _lib.stream())

        def fc():
            torch.addmm(b, x, w.t(), out=out)

        for name, fn in (("3xTF32 tcgen05", f3), ("fp32 SIMT", fs), ("cuBLAS fp32", fc)):
            try:
                fn()
                torch.cuda.synchronize()
                err = float((out[:4096] - ref).abs().max() / ref.abs().max())
                ms = timeit(fn, flush)
                print(f"M={M:7d} N={N:4d} K={Kd:4d} bt={bt} {name:15s} {ms * 1e3:9.1f} us {2.0 * M * N * Kd / ms / 1e9:8.2f} TFLOP/s"
                      f"  {(M * Kd + M * N) * 4 / ms / 1e6:7.1f} GB/s  rel err {err:.2e}", flush=True)
            except Exception as e:  # noqa: BLE001
                print(f"M={M} N={N} K={Kd} {name}: {e}", flush=True)


if __name__ == "__main__":
    main()

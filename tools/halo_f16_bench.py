"""Times the halo-strip convolution with fp16 activation storage (kind::f16) against the TF32 / fp32-storage variant on
custom_resnet18's stage-1 / stage-2 shapes, and the cluster GroupNorm with fp16 vs fp32 storage (CUDA events, median of
7, L2 flushed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200 import _lib
from avlen_b200 import nn as K


def timeit(fn, flush, n=7):
    ts = []
    for _ in range(n):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[n // 2]


def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 4800
    lib = _lib.lib()
    if len(sys.argv) > 2:
        K.set_conv_halo(1, int(sys.argv[2]))   # strip height (output rows per strip)
        print("strip rows", sys.argv[2])
    tma = int(os.environ.get("AVL_HALO_TMA", "1"))
    lib.avl_set_tc_conv_halo_tma(tma)
    st256 = int(os.environ.get("AVL_HALO_ST256", "1"))
    lib.avl_set_wide_stores(st256)
    print(f"-- halo strips by {'TMA' if tma else 'cp.async'}, {'32' if st256 else '16'}-byte stores")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    for name, H, W, C, Co, k in (("layer1 3x3 16->16 @64", 64, 64, 16, 16, 3), ("layer2 3x3 32->32 @32", 32, 32, 32, 32, 3),
                                 ("layer3 3x3 64->64 @16", 16, 16, 64, 64, 3), ("conv1 7x7 4->16 @64", 64, 64, 4, 16, 7)):
        x32 = torch.randn(B, H, W, C, device="cuda")
        w32 = K.round_to_tf32((torch.randn(Co, k, k, C, device="cuda") / (C * k * k) ** 0.5).contiguous())
        for in16, out16 in ((0, 0), (0, 1), (1, 1)):
            if in16 and C % 16:
                continue
            x, w = (x32.half(), w32.half()) if in16 else (x32, w32)
            y = torch.empty(B, H, W, Co, device="cuda", dtype=torch.float16 if out16 else torch.float32)

            def f():
                return lib.avl_tc_conv_halo_f16(x.data_ptr(), in16, B, H, W, C, w.data_ptr(), Co, k, k, k // 2, 0, y.data_ptr(),
                                                out16, _lib.stream())
            if f() != 0:  # -2: shape not covered by the halo-strip kernel (served by the TMA im2col kernel)
                continue
            ms = timeit(f, flush)
            byts = x.numel() * x.element_size() + y.numel() * y.element_size()
            print(f"B={B} {name:24s} in={'f16' if in16 else 'f32'} out={'f16' if out16 else 'f32'} {ms * 1e3:9.1f} us "
                  f"{byts / ms / 1e6:8.1f} GB/s  {2.0 * B * H * W * Co * C * k * k / ms / 1e9:7.2f} TFLOP/s", flush=True)
    for name, HW, C in (("GN 64x64x16", 4096, 16), ("GN 32x32x32", 1024, 32)):
        g, b = torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        x32 = torch.randn(B, HW, C, device="cuda")
        r32 = torch.randn(B, HW, C, device="cuda")
        for f16 in (0, 1):
            x, r = (x32.half(), r32.half()) if f16 else (x32, r32)

            def f():
                if f16:
                    rc = lib.avl_groupnorm_fwd_cluster_f16(x.data_ptr(), g.data_ptr(), b.data_ptr(), r.data_ptr(), x.data_ptr(), 1,
                                                           B, HW, C, 16, 1e-5, 1, _lib.stream())
                else:
                    rc = lib.avl_groupnorm_fwd_cluster(x.data_ptr(), g.data_ptr(), b.data_ptr(), r.data_ptr(), x.data_ptr(), B, HW,
                                                       C, 16, 1e-5, 1, _lib.stream())
                assert rc == 0, rc
            f()
            ms = timeit(f, flush)
            byts = 3 * x.numel() * x.element_size()
            print(f"B={B} {name:24s} storage={'f16' if f16 else 'f32'} (+residual, in place) {ms * 1e3:9.1f} us {byts / ms / 1e6:8.1f} GB/s",
                  flush=True)


if __name__ == "__main__":
    main()

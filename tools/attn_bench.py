"""Times the varlen self-attention kernels (avl_attn_self_fwd / bwd, head dim 32, 8 heads) on a PPO-minibatch-sized
problem: B samples with 40..151 valid tokens each (CUDA events, median of 7, L2 flushed)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200 import _lib
from avlen_b200 import nn as K  # noqa: F401


def main():
    B, D = (int(sys.argv[1]) if len(sys.argv) > 1 else 4800), 256
    g = torch.Generator().manual_seed(0)
    lens = torch.randint(40, 152, (B,), generator=g)
    off = torch.zeros(B + 1, dtype=torch.int32)
    off[1:] = torch.cumsum(lens, 0)
    R = int(off[-1])
    qkv = torch.randn(R, 3 * D, generator=g).cuda()
    dout = torch.randn(R, D, generator=g).cuda()
    off = off.cuda()
    out, lse = torch.empty(R, D, device="cuda"), torch.empty(R, D // 32, device="cuda")
    dqkv = torch.empty(R, 3 * D, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    pair_flops = float((lens.double() ** 2).sum()) * (D // 32)

    def fwd():
        _lib.call("avl_attn_self_fwd", qkv.data_ptr(), off.data_ptr(), B, D, out.data_ptr(), lse.data_ptr(), _lib.stream())

    def bwd():
        _lib.call("avl_attn_self_bwd", qkv.data_ptr(), off.data_ptr(), B, D, out.data_ptr(), lse.data_ptr(), dout.data_ptr(),
                  dqkv.data_ptr(), _lib.stream())

    mode = int(os.environ.get("AVL_ATTN_TC", "1"))
    _lib.lib().avl_set_attn_tc(mode)
    print(f"-- attn_tc={mode} (1: 3xTF32 warp MMAs, 0: register-tiled fp32)")
    fwd()
    # reference check on a few samples (fp64)
    for b in (0, B // 2, B - 1):
        r0, r1 = int(off[b]), int(off[b + 1])
        q, k, v = qkv[r0:r1].double().view(r1 - r0, 3, D // 32, 32).unbind(1)
        p = torch.softmax(torch.einsum("ihd,jhd->hij", q, k) / 32 ** 0.5, -1)
        ref = torch.einsum("hij,jhd->ihd", p, v).reshape(r1 - r0, D)
        print(f"sample {b}: V={r1 - r0} fwd max err {float((out[r0:r1].double() - ref).abs().max()):.2e}")
    for name, fn, fl in (("fwd", fwd, 2 * 64 * pair_flops), ("bwd", bwd, 2 * 224 * pair_flops)):
        ts = []
        for _ in range(7):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ms = sorted(ts)[3]
        print(f"attn_self_{name}: B={B} rows={R} {ms * 1e3:9.1f} us  {fl / ms / 1e9:7.2f} TFLOP/s (fp32 FMA)", flush=True)


if __name__ == "__main__":
    main()

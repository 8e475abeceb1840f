"""Launches ONE kernel of interest a few times (for `ncu -k regex:<name> -s 2 -c 1`); diagnostic.
    python tools/ncu_probe.py wgrad_l1 | wgrad_l3 | gn_bwd_l1 | halo_s2 | halo_f16_l1 | resize_u8 | tma_conv_l4[_b64] |
    attn_tc_fwd | audio_render | audio_spec | audio_spectral"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch

from avlen_b200 import nn as K


def main():
    what = sys.argv[1]
    B = int(sys.argv[2]) if len(sys.argv) > 2 else 4800
    K.set_tensor_cores(1)
    if what.startswith("wgrad"):
        H, C, Co = {"wgrad_l1": (64, 16, 16), "wgrad_l2": (32, 32, 32), "wgrad_l3": (16, 64, 64), "wgrad_l4": (8, 128, 128)}[what]
        x = torch.randn(B, H, H, C, device="cuda")
        gy = torch.randn(B, H, H, Co, device="cuda")
        for _ in range(4):
            K.conv2d_wgrad_tc(x, gy, (Co, C, 3, 3), 1, 1)
    elif what == "gn_bwd_l1":
        x = torch.randn(B, 64, 64, 16, device="cuda", requires_grad=True)
        g = torch.ones(16, device="cuda", requires_grad=True)
        b = torch.zeros(16, device="cuda", requires_grad=True)
        gy = torch.randn(B, 64, 64, 16, device="cuda")
        y = K.groupnorm(x, g, b, 16, 1e-5, relu=True)
        for _ in range(4):
            y.backward(gy, retain_graph=True)
    elif what == "halo_s2":
        x = torch.randn(B, 64, 64, 16, device="cuda")
        w = torch.randn(32, 16, 3, 3, device="cuda") / 12
        for _ in range(4):
            K._conv2d_raw(x, w, None, 2, 1)
    elif what == "resize_u8":
        x = torch.randint(0, 256, (B, 128, 128, 3), dtype=torch.uint8, device="cuda")
        for _ in range(4):
            K.resize_half(x, 1 / 255.0, 4)
    elif what in ("tma_conv_l4", "tma_conv_l4_b64"):
        b = 64 if what.endswith("b64") else B
        x = torch.randn(b, 8, 8, 128, device="cuda")
        w = torch.randn(128, 128, 3, 3, device="cuda") / 34
        for _ in range(4):
            K._conv2d_raw(x, w, None, 1, 1)
    elif what in ("attn_tc_fwd", "attn_tc_bwd"):
        from avlen_b200 import _lib
        g = torch.Generator().manual_seed(0)
        lens = torch.randint(40, 152, (B,), generator=g)
        off = torch.zeros(B + 1, dtype=torch.int32)
        off[1:] = torch.cumsum(lens, 0)
        R, D = int(off[-1]), 256
        qkv, dout = torch.randn(R, 3 * D, generator=g).cuda(), torch.randn(R, D, generator=g).cuda()
        off = off.cuda()
        out, lse, dqkv = torch.empty(R, D, device="cuda"), torch.empty(R, 8, device="cuda"), torch.empty(R, 3 * D, device="cuda")
        for _ in range(4):
            _lib.call("avl_attn_self_fwd", qkv.data_ptr(), off.data_ptr(), B, D, out.data_ptr(), lse.data_ptr(), _lib.stream())
            _lib.call("avl_attn_self_bwd", qkv.data_ptr(), off.data_ptr(), B, D, out.data_ptr(), lse.data_ptr(), dout.data_ptr(),
                      dqkv.data_ptr(), _lib.stream())
    elif what in ("audio_render", "audio_spec"):
        import numpy as np
        from avlen_b200 import synth
        from avlen_b200.audio import AudioRenderer
        n = 1024
        b = synth.make_audio_batch(1, n, fixed_len=16000, silent_frac=0.0, max_seconds=6)
        r = AudioRenderer(b["sr"])
        d = {k: torch.from_numpy(v).cuda() for k, v in b.items() if isinstance(v, np.ndarray)}
        for _ in range(4):
            if what == "audio_render":
                r.render(d["sounds"], d["clip_off"], d["index"], d["rirs"], d["rir_off"], d["rir_len"], d["silent"],
                         want_audiogoal=False)
            else:
                r.compute_spectrogram(torch.randn(n, 2, b["sr"], device="cuda"))
    elif what == "audio_spectral":
        import numpy as np
        from avlen_b200 import synth
        from avlen_b200.audio import AudioRenderer, SpectralSoundBank
        n = 1024
        b = synth.make_audio_batch(1, n, fixed_len=16000, silent_frac=0.0, max_seconds=6)
        r = AudioRenderer(b["sr"])
        d = {k: torch.from_numpy(v).cuda() for k, v in b.items() if isinstance(v, np.ndarray)}
        rs = r.rir_spectra(d["rirs"], d["rir_off"], d["rir_len"])
        sb = SpectralSoundBank(r, d["sounds"], b["clip_off_all"], b["clip_len_all"])
        row = torch.arange(n, device="cuda", dtype=torch.int64)
        for _ in range(4):
            r.render_spectral(sb.spectra, sb.rows(b["clip_id"]), d["index"], rs, row, d["silent"], want_audiogoal=False)
    elif what == "halo_f16_l1":
        from avlen_b200 import _lib
        x = torch.randn(B, 64, 64, 16, device="cuda").half()
        w = (torch.randn(16, 3, 3, 16, device="cuda") / 12).half()
        y = torch.empty(B, 64, 64, 16, device="cuda", dtype=torch.float16)
        for _ in range(4):
            rc = _lib.lib().avl_tc_conv_halo_f16(x.data_ptr(), 1, B, 64, 64, 16, w.data_ptr(), 16, 3, 3, 1, 0, y.data_ptr(), 1,
                                                _lib.stream())
            assert rc == 0, rc
    torch.cuda.synchronize()
    print("ok")


if __name__ == "__main__":
    main()

"""Host-emulated run of the encoder kernels (csrc/conv.cu: im2col-gather GEMM with OIHW weights, GroupNorm,
resize, pooling) against PyTorch."""
import ctypes

import pytest
import torch
import torch.nn.functional as F

vp, ci, ll, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


def _np(t):
    return t.detach().contiguous().numpy()


@pytest.mark.parametrize("shape", [(2, 9, 7, 2, 5, 5, 5, 2, 0), (2, 8, 8, 3, 16, 7, 7, 1, 3), (1, 10, 6, 16, 32, 3, 3, 2, 1),
                                   (2, 6, 6, 16, 32, 1, 1, 2, 0), (3, 4, 3, 8, 20, 4, 3, 1, 0),
                                   (2, 8, 8, 32, 10, 8, 8, 1, 0)])  # last: split-K path (K = 2048, one output tile)
def test_conv2d(emul_lib, shape):
    N, H, W, C, Co, KH, KW, s, p = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(N, H, W, C, generator=g)
    w = torch.randn(Co, C, KH, KW, generator=g)
    b, sc = torch.randn(Co, generator=g), torch.rand(Co, generator=g) + 0.5
    y0 = F.conv2d(x.permute(0, 3, 1, 2), w, None, s, p)
    res = torch.randn(*y0.permute(0, 2, 3, 1).shape, generator=g)
    ref = F.relu(y0 * sc.view(1, -1, 1, 1) + b.view(1, -1, 1, 1) + res.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    out = torch.zeros(ref.shape).contiguous()
    emul_lib.avl_conv2d_fwd.argtypes = [vp, ci, ci, ci, ci, vp, ci, ci, ci, ci, ci, vp, vp, vp, ll, ci, vp, ll, vp]
    xn, wn, bn, sn, rn, on = _np(x), _np(w), _np(b), _np(sc), _np(res), out.numpy()
    rc = emul_lib.avl_conv2d_fwd(xn.ctypes.data, N, H, W, C, wn.ctypes.data, Co, KH, KW, s, p, sn.ctypes.data,
                                 bn.ctypes.data, rn.ctypes.data, Co, 1, on.ctypes.data, Co, None)
    assert rc == 0
    assert (out - ref).abs().max() < 1e-4 * max(1.0, float(ref.abs().max()))


@pytest.mark.parametrize("C,H,W", [(16, 6, 5), (32, 4, 4), (128, 3, 2)])
def test_groupnorm(emul_lib, C, H, W):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(2, H, W, C, generator=g) * 2 + 1
    ga, be = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    res = torch.randn(2, H, W, C, generator=g)
    ref = F.relu(F.group_norm(x.permute(0, 3, 1, 2), 16, ga, be, 1e-5) + res.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    out = torch.zeros_like(x)
    emul_lib.avl_groupnorm_fwd.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, cf, ci, vp]
    xn, gn, bn, rn, on = _np(x), _np(ga), _np(be), _np(res), out.numpy()
    rc = emul_lib.avl_groupnorm_fwd(xn.ctypes.data, gn.ctypes.data, bn.ctypes.data, rn.ctypes.data, on.ctypes.data, 2,
                                    H * W, C, 16, 1e-5, 1, None)
    assert rc == 0
    assert (out - ref).abs().max() < 1e-4


@pytest.mark.parametrize("C,H,W", [(16, 12, 11), (64, 5, 4)])
def test_groupnorm_split(emul_lib, C, H, W):
    import numpy as np
    g = torch.Generator().manual_seed(C + 1)
    x = torch.randn(3, H, W, C, generator=g) * 2 + 1
    ga, be = torch.rand(C, generator=g) + 0.5, torch.randn(C, generator=g)
    res = torch.randn(3, H, W, C, generator=g)
    ref = F.relu(F.group_norm(x.permute(0, 3, 1, 2), 16, ga, be, 1e-5) + res.permute(0, 3, 1, 2)).permute(0, 2, 3, 1)
    out = torch.zeros_like(x)
    st = np.zeros(3 * 16 * 2 + 3 * C, np.float64)
    emul_lib.avl_groupnorm_fwd_split.argtypes = [vp, vp, vp, vp, vp, ci, ci, ci, ci, cf, ci, vp, vp]
    xn, gn, bn, rn, on = _np(x), _np(ga), _np(be), _np(res), out.numpy()
    rc = emul_lib.avl_groupnorm_fwd_split(xn.ctypes.data, gn.ctypes.data, bn.ctypes.data, rn.ctypes.data,
                                          on.ctypes.data, 3, H * W, C, 16, 1e-5, 1, st.ctypes.data, None)
    assert rc == 0
    assert (out - ref).abs().max() < 1e-4

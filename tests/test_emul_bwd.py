"""Host-emulated run of the backward kernels (csrc/nn_bwd.cu: conv dgrad / wgrad, GroupNorm backward, GRU
forward + backward through time) against PyTorch autograd."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

vp, ci, ll, cf = ctypes.c_void_p, ctypes.c_int, ctypes.c_longlong, ctypes.c_float


def _np(t):
    return t.detach().contiguous().numpy()


def _ptr(a):
    return None if a is None else a.ctypes.data


@pytest.mark.parametrize("shape", [(2, 9, 7, 2, 5, 5, 5, 2, 0), (2, 8, 8, 3, 16, 7, 7, 1, 3), (1, 10, 6, 16, 32, 3, 3, 2, 1),
                                   (2, 6, 6, 16, 32, 1, 1, 2, 0), (3, 4, 3, 8, 20, 4, 3, 1, 0), (2, 12, 12, 4, 8, 8, 8, 4, 0),
                                   (1, 5, 5, 70, 130, 3, 3, 1, 1)])
def test_conv_dgrad_wgrad(emul_lib, shape):
    N, H, W, C, Co, KH, KW, s, p = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(N, H, W, C, generator=g, requires_grad=True)
    w = torch.randn(Co, C, KH, KW, generator=g, requires_grad=True)
    b = torch.randn(Co, generator=g, requires_grad=True)
    y = F.conv2d(x.permute(0, 3, 1, 2), w, b, s, p).permute(0, 2, 3, 1)
    gy = torch.randn(*y.shape, generator=g)
    y.backward(gy)
    emul_lib.avl_conv2d_dgrad.argtypes = [vp, vp, vp] + [ci] * 10 + [vp]
    emul_lib.avl_conv2d_wgrad.argtypes = [vp, vp, vp, vp] + [ci] * 9 + [vp]
    gyn, wn, xn = _np(gy), _np(w), _np(x)
    dx = np.full(x.shape, 0.5, np.float32)
    assert emul_lib.avl_conv2d_dgrad(_ptr(gyn), _ptr(wn), _ptr(dx), N, H, W, C, Co, KH, KW, s, p, 1, None) == 0
    assert np.abs(dx - 0.5 - _np(x.grad)).max() < 2e-4 * max(1.0, float(x.grad.abs().max()))
    assert emul_lib.avl_conv2d_dgrad(_ptr(gyn), _ptr(wn), _ptr(dx), N, H, W, C, Co, KH, KW, s, p, 0, None) == 0
    assert np.abs(dx - _np(x.grad)).max() < 2e-4 * max(1.0, float(x.grad.abs().max()))
    dw = np.zeros(w.shape, np.float32)
    db = np.zeros(Co, np.float32)
    assert emul_lib.avl_conv2d_wgrad(_ptr(xn), _ptr(gyn), _ptr(dw), _ptr(db), N, H, W, C, Co, KH, KW, s, p, None) == 0
    assert np.abs(dw - _np(w.grad)).max() < 2e-4 * max(1.0, float(w.grad.abs().max()))
    assert np.abs(db - _np(b.grad)).max() < 2e-4 * max(1.0, float(b.grad.abs().max()))


@pytest.mark.parametrize("C,H,W,relu,res", [(16, 6, 5, 1, 1), (32, 4, 4, 1, 0), (128, 3, 2, 0, 1), (64, 5, 3, 0, 0)])
def test_groupnorm_bwd(emul_lib, C, H, W, relu, res):
    g = torch.Generator().manual_seed(C + relu)
    x = (torch.randn(2, H, W, C, generator=g) * 2 + 1).requires_grad_()
    ga = (torch.rand(C, generator=g) + 0.5).requires_grad_()
    be = torch.randn(C, generator=g).requires_grad_()
    r = torch.randn(2, H, W, C, generator=g).requires_grad_()
    y = F.group_norm(x.permute(0, 3, 1, 2), 16, ga, be, 1e-5).permute(0, 2, 3, 1)
    if res:
        y = y + r
    if relu:
        y = F.relu(y)
    gy = torch.randn(*y.shape, generator=g)
    y.backward(gy)
    emul_lib.avl_groupnorm_bwd.argtypes = [vp] * 8 + [ci] * 4 + [cf, ci, vp]
    xn, yn, gyn, gn = _np(x), _np(y), _np(gy), _np(ga)
    dx, dres = np.zeros(x.shape, np.float32), np.zeros(x.shape, np.float32)
    dg, db = np.zeros(C, np.float32), np.zeros(C, np.float32)
    rc = emul_lib.avl_groupnorm_bwd(_ptr(xn), _ptr(yn), _ptr(gyn), _ptr(gn), _ptr(dx), _ptr(dres) if res else None,
                                    _ptr(dg), _ptr(db), 2, H * W, C, 16, 1e-5, relu, None)
    assert rc == 0
    assert np.abs(dx - _np(x.grad)).max() < 2e-4
    assert np.abs(dg - _np(ga.grad)).max() < 2e-4 * max(1.0, float(ga.grad.abs().max()))
    assert np.abs(db - _np(be.grad)).max() < 2e-4 * max(1.0, float(be.grad.abs().max()))
    if res:
        assert np.abs(dres - _np(r.grad)).max() < 1e-6


@pytest.mark.parametrize("T,N,I,H", [(1, 3, 20, 16), (6, 4, 24, 32), (5, 2, 70, 40)])
def test_gru_forward_backward(emul_lib, T, N, I, H):
    """nn.GRU stepped with h * mask_t (rnn_state_encoder.py:84-87, :124-136) vs the fused kernels."""
    g = torch.Generator().manual_seed(T * 100 + N)
    gru = torch.nn.GRU(I, H)
    for q in gru.parameters():
        q.data = torch.randn(q.shape, generator=g) * 0.3
    x = torch.randn(T * N, I, generator=g, requires_grad=True)
    h0 = torch.randn(N, H, generator=g, requires_grad=True)
    masks = (torch.rand(T * N, generator=g) > 0.3).float()
    h = h0.unsqueeze(0)
    outs = []
    for t in range(T):
        o, h = gru(x[t * N:(t + 1) * N].unsqueeze(0), h * masks[t * N:(t + 1) * N].view(1, N, 1))
        outs.append(o[0])
    out = torch.cat(outs)
    g_out = torch.randn(T * N, H, generator=g)
    g_last = torch.randn(N, H, generator=g)
    (out * g_out).sum().add((h[0] * g_last).sum()).backward()

    emul_lib.avl_gru_workspace_bytes.restype = ll
    emul_lib.avl_gru_workspace_bytes.argtypes = [ci] * 5
    emul_lib.avl_gru_forward.argtypes = [ci] * 4 + [vp] * 10 + [ci, vp]
    emul_lib.avl_gru_backward.argtypes = [ci] * 4 + [vp] * 14
    ws = np.zeros(int(emul_lib.avl_gru_workspace_bytes(T, N, I, H, 1)) // 4 + 64, np.float32)
    xn, hn, mn = _np(x), _np(h0), _np(masks)
    wih, whh, bih, bhh = (_np(q) for q in (gru.weight_ih_l0, gru.weight_hh_l0, gru.bias_ih_l0, gru.bias_hh_l0))
    o = np.zeros((T * N, H), np.float32)
    hl = np.zeros((N, H), np.float32)
    assert emul_lib.avl_gru_forward(T, N, I, H, _ptr(xn), _ptr(hn), _ptr(mn), _ptr(wih), _ptr(whh), _ptr(bih), _ptr(bhh),
                                    _ptr(o), _ptr(hl), _ptr(ws), 1, None) == 0
    assert np.abs(o - _np(out)).max() < 1e-5
    assert np.abs(hl - _np(h[0])).max() < 1e-5
    # inference layout (no saved state) gives the same result
    ws2 = np.zeros(int(emul_lib.avl_gru_workspace_bytes(T, N, I, H, 0)) // 4 + 64, np.float32)
    o2 = np.zeros_like(o)
    assert emul_lib.avl_gru_forward(T, N, I, H, _ptr(xn), _ptr(hn), _ptr(mn), _ptr(wih), _ptr(whh), _ptr(bih), _ptr(bhh),
                                    _ptr(o2), None, _ptr(ws2), 0, None) == 0
    assert np.array_equal(o, o2)
    dx, dh0 = np.zeros_like(xn), np.zeros_like(hn)
    dwih, dwhh = np.zeros_like(wih), np.zeros_like(whh)
    dbih, dbhh = np.zeros_like(bih), np.zeros_like(bhh)
    gon, gln = _np(g_out), _np(g_last)
    assert emul_lib.avl_gru_backward(T, N, I, H, _ptr(xn), _ptr(mn), _ptr(wih), _ptr(whh), _ptr(gon), _ptr(gln),
                                     _ptr(dx), _ptr(dh0), _ptr(dwih), _ptr(dwhh), _ptr(dbih), _ptr(dbhh), _ptr(ws),
                                     None) == 0
    for a, b in ((dx, x.grad), (dh0, h0.grad), (dwih, gru.weight_ih_l0.grad), (dwhh, gru.weight_hh_l0.grad),
                 (dbih, gru.bias_ih_l0.grad), (dbhh, gru.bias_hh_l0.grad)):
        assert np.abs(a - _np(b)).max() < 2e-4 * max(1.0, float(b.abs().max()))

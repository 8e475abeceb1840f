"""Rows K and L in host emulation: the dialog state encoder (avl_dialog_forward / backward, csrc/smt.cu) and the CLIP
text tower (avl_clip_text_forward, csrc/clip_text.cu) run on the CPU from the same kernel source against the oracle."""
import ctypes

import numpy as np
import pytest
import torch

from avlen_b200.savi.models.clip_text import clip_param_keys
from avlen_b200.savi.models.dialog_state_encoder import DIALOG_PARAM_KEYS
from oracle import clip_torch
from oracle import models_torch as OM

vp, ci = ctypes.c_void_p, ctypes.c_int


@pytest.mark.parametrize("with_dialog,indexed", [(True, False), (False, False), (True, True)])
def test_dialog_encoder_forward_backward_matches_oracle(emul_lib, with_dialog, indexed):
    lib = emul_lib
    lib.avl_dialog_workspace_bytes.restype = ctypes.c_longlong
    lib.avl_dialog_workspace_bytes.argtypes = [ci] * 4
    lib.avl_dialog_forward.argtypes = [ci, ci, ci, vp, vp, ci, vp, vp, vp, vp, vp, ci, vp, vp, vp, vp, ci, vp]
    lib.avl_dialog_backward.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, vp, vp, vp, vp, vp]
    g = torch.Generator().manual_seed(5)
    B, K, D = 4, 3, 256
    enc = OM.DialogStateEncoder(2 * D, dim_feedforward=D)
    enc.load_state_dict(OM.seeded_state_dict(enc, 4))
    n_mem = 6 if indexed else B
    env_index = torch.tensor([5, 0, 2, 5], dtype=torch.int32) if indexed else None
    x = torch.randn(B, D, generator=g)
    mem = torch.randn(K, n_mem, D, generator=g)
    masks = (torch.rand(B, K, generator=g) > 0.4).float()
    masks[1] = 0
    d_emb = torch.randn(B, D, generator=g) if with_dialog else None
    step = torch.tensor([0, 2, 1, 99], dtype=torch.int32)
    goal = torch.randn(B, D, generator=g)
    gout = torch.randn(B, D, generator=g)

    xr = x.clone().requires_grad_(True)
    dr = d_emb.clone().requires_grad_(True) if with_dialog else None
    gr = goal.clone().requires_grad_(True)
    mem_b = mem[:, env_index.long()] if indexed else mem
    out_ref = enc(xr, mem_b, masks, dr, step, goal=gr)
    (out_ref * gout).sum().backward()
    sd = dict(enc.named_parameters())
    params = [sd[k].detach().contiguous().numpy() for k in DIALOG_PARAM_KEYS]
    grads = [np.zeros_like(p) for p in params]
    pe = enc.pos_encode.pe[:, 0].contiguous().numpy()

    nbytes = lib.avl_dialog_workspace_bytes(B, K, D, 1)
    ws = np.zeros(nbytes, np.uint8)
    out = np.zeros((B, D), np.float32)
    ptab = (vp * len(params))(*[p.ctypes.data for p in params])
    gtab = (vp * len(params))(*[q.ctypes.data for q in grads])
    xn, memn, mn, gn, gon, sn = (x.numpy().copy(), mem.numpy().copy(), masks.numpy().copy(), goal.numpy().copy(),
                                 gout.numpy().copy(), step.numpy().copy())
    dn = d_emb.numpy().copy() if with_dialog else None
    ein = env_index.numpy().copy() if indexed else None
    rc = lib.avl_dialog_forward(B, K, D, xn.ctypes.data, memn.ctypes.data, n_mem,
                                None if ein is None else ein.ctypes.data, mn.ctypes.data,
                                None if dn is None else dn.ctypes.data, sn.ctypes.data, pe.ctypes.data, pe.shape[0],
                                gn.ctypes.data, ctypes.cast(ptab, vp), out.ctypes.data, ws.ctypes.data, 1, None)
    assert rc == 0
    ref = out_ref.detach().numpy()
    assert np.abs(out - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())

    dx, dd, dgoal = (np.zeros((B, D), np.float32) for _ in range(3))
    rc = lib.avl_dialog_backward(B, K, D, int(with_dialog), gn.ctypes.data, ctypes.cast(ptab, vp), ctypes.cast(gtab, vp),
                                 gon.ctypes.data, dx.ctypes.data, dd.ctypes.data if with_dialog else None,
                                 dgoal.ctypes.data, ws.ctypes.data, None)
    assert rc == 0
    for k, gk in zip(DIALOG_PARAM_KEYS, grads):
        want = sd[k].grad
        want = np.zeros_like(gk) if want is None else want.numpy()
        err = np.abs(gk - want).max()
        assert err < 1e-4 * max(1.0, np.abs(want).max()), (k, err)
    for got, want in ((dx, xr.grad), (dgoal, gr.grad)) + (((dd, dr.grad),) if with_dialog else ()):
        want = want.numpy()
        assert np.abs(got - want).max() < 1e-4 * max(1.0, np.abs(want).max())


@pytest.mark.parametrize("dedupe", [1, 0])
def test_clip_text_tower_matches_oracle(emul_lib, dedupe):
    lib = emul_lib
    lib.avl_clip_text_workspace_bytes.restype = ctypes.c_longlong
    lib.avl_clip_text_workspace_bytes.argtypes = [ci, ci]
    lib.avl_clip_text_forward.argtypes = [ci, ci, ci, ci, vp, vp, vp, vp, ci, vp]
    lib.avl_clip_text_status.argtypes = [ci, ci, vp, ctypes.POINTER(ci), ctypes.POINTER(ci)]
    torch.manual_seed(0)
    layers, L, vocab, B = 2, 12, 400, 5
    o = clip_torch.CLIPText(layers=layers, context=L, vocab=vocab).eval()
    o.load_state_dict(OM.seeded_state_dict(o, 8))
    tokens = torch.zeros(B, L, dtype=torch.long)
    for b, n in ((0, 4), (2, 9), (3, 1)):  # rows 1 and 4 stay all-zero (no active query)
        tokens[b, 0] = vocab - 2
        tokens[b, 1:1 + n] = torch.randint(1, vocab - 2, (n,))
        tokens[b, 1 + n] = vocab - 1
    with torch.no_grad():
        ref = o.encode_text(tokens).numpy()
    assert lib.avl_clip_text_param_count(layers) == len(clip_param_keys(layers))
    sd = dict(o.named_parameters())
    params = [sd[k].detach().contiguous().numpy() for k in clip_param_keys(layers)]
    ptab = (vp * len(params))(*[p.ctypes.data for p in params])
    ws = np.zeros(lib.avl_clip_text_workspace_bytes(B, L), np.uint8)
    out = np.zeros((B, 512), np.float32)
    tn = tokens.numpy().copy()
    rc = lib.avl_clip_text_forward(B, L, vocab, layers, tn.ctypes.data, ctypes.cast(ptab, vp), out.ctypes.data,
                                   ws.ctypes.data, dedupe, None)
    assert rc == 0
    assert np.abs(out - ref).max() < 2e-5 * max(1.0, np.abs(ref).max())
    a, b = ci(0), ci(0)
    assert lib.avl_clip_text_status(B, L, ws.ctypes.data, ctypes.byref(a), ctypes.byref(b)) == 0
    assert (a.value, b.value) == ((4, 4 * L) if dedupe else (B, B * L))  # 3 active rows + 1 shared all-zero row
    np.testing.assert_array_equal(out[1], out[4])

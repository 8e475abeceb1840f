"""Functional wrappers over the encoder kernels (csrc/conv.cu): NHWC convolution / FC, GroupNorm(+residual+ReLU),
area resize, pooling.  CUDA tensors in, CUDA tensors out, no CPU path."""
from __future__ import annotations

from ctypes import c_float, c_int, c_longlong, c_void_p

import torch

from . import _lib
from ._lib import call, dptr, fptr, stream

P, I, L, F = c_void_p, c_int, c_longlong, c_float
_lib.register({
    "avl_conv2d_fwd": [P, I, I, I, I, P, I, I, I, I, I, P, P, P, L, I, P, L, P],
    "avl_groupnorm_fwd": [P, P, P, P, P, I, I, I, I, F, I, P],
    "avl_groupnorm_fwd_split": [P, P, P, P, P, I, I, I, I, F, I, P, P],
    "avl_groupnorm_fwd_cluster": [P, P, P, P, P, I, I, I, I, F, I, P],
    "avl_resize_half": [P, P, I, I, I, I, I, F, P],
    "avl_pad_channels": [P, P, L, I, I, P],
    "avl_concat_rgbd": [P, P, P, L, I, I, F, P],
    "avl_append_planes": [P, P, P, I, I, I, I, P],
    "avl_maxpool3x3s2": [P, P, I, I, I, I, P],
    "avl_avgpool_global": [P, P, I, I, I, P],
    "avl_onehot_linear": [P, P, P, P, L, I, I, I, P],
    "avl_copy_cols": [P, L, P, L, I, I, P],
    "avl_gemm": [P, L, L, P, L, L, P, L, I, I, I, P, I, I, I, P],
    "avl_layernorm_fwd": [P, P, P, P, P, P, I, I, P],
    "avl_layernorm_bwd": [P, P, P, P, P, P, P, P, I, I, P],
    "avl_attn_self_fwd": [P, P, I, I, P, P, P],
    "avl_attn_self_bwd": [P, P, I, I, P, P, P, P, P],
    "avl_attn_cross_fwd": [P, P, P, I, I, P, P, P],
    "avl_attn_cross_bwd": [P, P, P, P, P, I, I, P, P, P],
    "avl_tc_gemm": [P, L, P, P, L, I, I, I, P, P, P, L, I, P, P],
    "avl_tc_conv2d_fwd": [P, I, I, I, I, P, I, I, I, I, I, P, P, P, L, I, P, L, P],
    "avl_set_tensor_cores": [I],
    "avl_get_tensor_cores": [],
    "avl_set_tc_conv_l1": [I],
    "avl_set_attn_tc": [I],
    "avl_set_tc_conv_tma": [I],
    "avl_set_wide_stores": [I],
    "avl_set_attn_qsplit": [I],
    "avl_set_tc_conv_halo_small_grid": [I],
    "avl_set_tc_splitk_fill": [I],
    "avl_tc_conv2d_dgrad_s2": [P, I, I, I, I, P, I, P, P],
    "avl_set_pdl": [I],
    "avl_set_tc_conv_halo_tma": [I],
    "avl_tc_conv_tma_count": [],
    "avl_set_tc_splitk": [I],
    "avl_set_tc_splitk_cluster": [I],
    "avl_set_tc_stages": [I],
    "avl_set_tc_swizzle": [I],
    "avl_set_tc_tma": [I],
    "avl_set_tc_3xtf32": [I],
    "avl_tc_gemm_3x": [P, L, P, L, I, P, L, I, I, I, P, P, L, I, P, P],
    "avl_tc_wgrad_3x": [P, L, P, L, P, L, I, I, I, P, P],
    "avl_set_f16_activations": [I],
    "avl_set_resnet_graphs": [I],
    "avl_set_tc_conv_halo_group": [I],
    "avl_set_tc_conv_halo_stride2": [I],
    "avl_tc_conv_halo_f16": [P, I, I, I, I, I, P, I, I, I, I, I, P, I, P],
    "avl_groupnorm_fwd_cluster_f16": [P, P, P, P, P, I, I, I, I, I, F, I, P],
    "avl_set_wgrad_desc": [I, I],
    "avl_resnet18_param_count": [],
    "avl_resnet18_workspace_bytes": [I, I, I, P],
    "avl_resnet18_forward": [P, I, I, I, I, P, F, P, P, L, I, P, P],
    "avl_resnet18_forward_pair": [P, P, I, I, I, I, I, P, P, F, P, P, P, P, L, L, I, P, P, P],
    "avl_set_tc_conv_halo": [I, I],
    "avl_conv2d_dgrad": [P, P, P, I, I, I, I, I, I, I, I, I, I, P],
    "avl_conv2d_wgrad": [P, P, P, P, I, I, I, I, I, I, I, I, I, P],
    "avl_relu_mask": [P, L, P, L, L, I, P],
    "avl_groupnorm_bwd": [P, P, P, P, P, P, P, P, I, I, I, I, F, I, P],
    "avl_groupnorm_bwd_cluster": [P, P, P, P, P, P, P, P, I, I, I, I, F, I, P, P],
    "avl_gru_workspace_bytes": [I, I, I, I, I],
    "avl_gru_forward": [I, I, I, I, P, P, P, P, P, P, P, P, P, P, I, P],
    "avl_gru_backward": [I, I, I, I, P, P, P, P, P, P, P, P, P, P, P, P, P, P],
    "avl_resnet_graph_stats": [I],
    "avl_f16_overflow": [I],
    "avl_resize_half_typed": [P, I, P, P, I, I, I, I, I, F, P],
    "avl_pack_conv_weight": [P, I, I, I, I, I, I, P, P],
    "avl_zero_upsample2": [P, P, I, I, I, I, I, I, P],
    "avl_tc_conv2d_wgrad_workspace": [I, I, I, I, I, I, I, I, I],
    "avl_tc_conv2d_wgrad": [P, P, P, I, I, I, I, I, I, I, I, I, I, P, L, P],
}, {"avl_gru_workspace_bytes": c_longlong, "avl_resnet18_workspace_bytes": c_longlong,
    "avl_resnet_graph_stats": c_longlong, "avl_tc_conv2d_wgrad_workspace": c_longlong})

_gn_scratch = {}
_gn_bwd_scratch = {}
_gn_cluster = [True]
_tc_min_rows = [64]


def set_groupnorm_cluster(on) -> bool:
    """Single-pass cluster GroupNorm (csrc/gn_cluster.cu) on / off (off: the two-pass kernels)."""
    old = _gn_cluster[0]
    _gn_cluster[0] = bool(on)
    return old


def set_tensor_cores(level) -> int:
    """tcgen05 (TF32) level: 0/False = fp32 SIMT only, 1/True = encoder convs + FCs (default), 2 = also the SMT
    dense layers.  Returns the previous level."""
    return int(_lib.lib().avl_set_tensor_cores(int(level)))


def set_conv_halo(on, rows=0) -> int:
    """Halo-strip tensor-core kernel for stride-1 same-padded convolutions (csrc/conv_halo_tc.cu) on / off."""
    return int(_lib.lib().avl_set_tc_conv_halo(int(on), int(rows)))


# ---- deferred joins: work enqueued on a side stream whose results are consumed later on another stream -------------
_pending_events = []


def defer_join(stream):
    """Record the completion of everything enqueued on ``stream`` so far; the consumer calls ``sync_pending()``."""
    ev = torch.cuda.Event()
    ev.record(stream)
    _pending_events.append(ev)


def sync_pending():
    """Make the current stream wait for every deferred producer (no host synchronisation)."""
    if _pending_events:
        cur = torch.cuda.current_stream()
        for ev in _pending_events:
            cur.wait_event(ev)
        _pending_events.clear()


def set_resnet_graphs(on: bool) -> bool:
    """CUDA-graph replay of repeated whole-network calls at small batch (csrc/resnet_fwd.cu).  Returns the old setting."""
    return bool(_lib.lib().avl_set_resnet_graphs(int(bool(on))))


def resnet_graph_stats():
    """(replays, captures) of whole-network CUDA graphs so far."""
    return int(_lib.lib().avl_resnet_graph_stats(0)), int(_lib.lib().avl_resnet_graph_stats(1))


def set_f16_activations(on: bool) -> bool:
    """fp16 storage of the fused GroupNorm ResNet-18's stem output and stage 1 (tensor-core path).  Returns the old
    setting."""
    return bool(_lib.lib().avl_set_f16_activations(int(bool(on))))


def f16_overflow(reset: bool = False) -> bool:
    """True if an fp16-stored activation saturated (|v| >= 65504) since the last reset.  Synchronises the device."""
    v = int(_lib.lib().avl_f16_overflow(int(bool(reset))))
    if v < 0:
        _lib.check(v, "avl_f16_overflow")
    return bool(v)


def check_f16_overflow() -> bool:
    """Call at a point where the host synchronises anyway (end of a PPO update): if an fp16-stored activation of the
    fused ResNet-18s saturated, switch the activation storage back to fp32 for the rest of the run and say so."""
    if not f16_overflow(reset=True):
        return False
    set_f16_activations(False)
    import warnings
    warnings.warn("avlen_b200: an fp16-stored encoder activation saturated (|v| >= 65504); activation storage falls "
                  "back to fp32 from here on (results since the last check used clamped values)", RuntimeWarning)
    return True


def tensor_cores_level() -> int:
    return int(_lib.lib().avl_get_tensor_cores())


def tensor_cores_enabled() -> bool:
    return tensor_cores_level() >= 1


def _packed_weight(w, c_pad=None, dgrad=False):
    """(Cout, C, KH, KW) -> (Cout, KH, KW, Cp) K-contiguous copy for the tensor-core path (input channels zero-padded
    to ``c_pad``), rounded to nearest onto the TF32 grid, written by ONE kernel (csrc/conv_bwd_tc.cu
    ``avl_pack_conv_weight``; no ATen permute / pad / bit-twiddling launches).  ``dgrad=True``: the weight of the
    data-gradient convolution instead, (C, KH, KW, Coutp) with both kernel axes flipped and the channel roles swapped
    (``c_pad`` then pads Cout).  Cached ON the owning parameter object per weight version (a cache keyed by address
    would hand a new module, allocated where a freed one lived, the old module's weights)."""
    owner = w._base if w._base is not None else w
    cache = getattr(owner, "_avl_packed", None)
    if cache is None:
        cache = {}
        try:
            owner._avl_packed = cache
        except AttributeError:
            pass
    key = (tuple(w.shape), c_pad, w.data_ptr(), bool(dgrad))
    hit = cache.get(key)
    if hit is not None and hit[0] == w._version:
        return hit[1]
    Cout, C, KH, KW = w.shape
    wc = w.detach()
    if not wc.is_contiguous():
        wc = wc.contiguous()
    cp = (Cout if dgrad else C) if c_pad is None else c_pad
    shape = (C, KH, KW, cp) if dgrad else (Cout, KH, KW, cp)
    # a stale entry is re-packed IN PLACE: the packed weight keeps its device address for the life of the parameter, so
    # a rollout step captured into a CUDA graph (which also captures this re-pack when it follows an optimizer step)
    # stays valid after the weights change — trainable encoders replay their step graphs like frozen ones
    pk = hit[1] if (hit is not None and tuple(hit[1].shape) == shape) else torch.empty(shape, device=w.device, dtype=torch.float32)
    call("avl_pack_conv_weight", fptr(wc), Cout, C, KH, KW, cp, int(bool(dgrad)), fptr(pk), stream())
    cache[key] = (w._version, pk)
    return pk


def round_to_tf32(t):
    """Round-to-nearest-even onto the TF32 grid (10 explicit mantissa bits), in place on a contiguous fp32 tensor.
    The tensor core TRUNCATES fp32 operands to TF32; weights that are already on the grid are consumed exactly, so the
    weight side of every tensor-core product carries a round-to-nearest (unbiased, half the magnitude) error."""
    bits = t.view(torch.int32)
    lsb = (bits >> 13) & 1
    bits.add_(0xFFF + lsb).bitwise_and_(~0x1FFF)  # Inf / NaN are not expected in weights
    return t


def pad_channels(x, c_out):
    """NHWC channel zero-padding (C -> c_out)."""
    N, H, W, C = x.shape
    y = torch.empty((N, H, W, c_out), device=x.device, dtype=torch.float32)
    call("avl_pad_channels", fptr(x), fptr(y), N * H * W, C, c_out, stream())
    return y



def conv_out(size, k, stride, pad):
    return (size + 2 * pad - k) // stride + 1


def _needs_grad(*ts):
    return torch.is_grad_enabled() and any(t is not None and t.requires_grad for t in ts)


def conv2d(x, w, bias=None, stride=1, pad=0, relu=False, scale=None, residual=None, out=None):
    """x (N,H,W,C) NHWC; w (Cout,C,KH,KW) OIHW (the reference's nn.Conv2d layout). Returns (N,OH,OW,Cout).
    Differentiable (custom backward on the dgrad / wgrad kernels) when an input requires grad."""
    if _needs_grad(x, w, bias):
        if scale is not None or residual is not None or out is not None:
            raise _lib.AvlenError("the differentiable conv2d takes no scale / residual / out arguments")
        return _ConvFn.apply(x, w, bias, int(stride), int(pad), bool(relu))
    return _conv2d_raw(x, w, bias, stride, pad, relu, scale, residual, out)


def _conv2d_raw(x, w, bias=None, stride=1, pad=0, relu=False, scale=None, residual=None, out=None):
    N, H, W, C = x.shape
    Cout, Cw, KH, KW = w.shape
    assert Cw == C or (Cw < C and C % 4 == 0), (Cw, C)  # x may carry zero-padded channels (tensor-core path)
    OH, OW = conv_out(H, KH, stride, pad), conv_out(W, KW, stride, pad)
    # >= 64 output positions: below a full 128-row tile the tensor-core kernel still wins through split-K
    tc = N * OH * OW >= _tc_min_rows[0] and x.data_ptr() % 16 == 0 and tensor_cores_enabled()
    if tc and C % 4 != 0:
        cp = (C + 3) // 4 * 4
        x, C = pad_channels(x, cp), cp
    if out is None:
        out = torch.empty((N, OH, OW, Cout), device=x.device, dtype=torch.float32)
        ldy = Cout
    else:  # (N*OH*OW, >=Cout) strided destination (column slice of a feature matrix)
        assert out.stride(-1) == 1
        ldy = out.stride(0)
    ldr = Cout if residual is not None else 0
    if tc:
        call("avl_tc_conv2d_fwd", fptr(x), N, H, W, C, fptr(_packed_weight(w, C)), Cout, KH, KW, stride, pad, fptr(scale),
             fptr(bias), fptr(residual), ldr, int(relu), out.data_ptr(), ldy, stream())
        return out
    if Cw != C:
        raise _lib.AvlenError("channel-padded input needs the tensor-core path")
    call("avl_conv2d_fwd", fptr(x), N, H, W, C, fptr(w), Cout, KH, KW, stride, pad, fptr(scale), fptr(bias),
         fptr(residual), ldr, int(relu), out.data_ptr(), ldy, stream())
    return out


def _conv2d_packed(x, pk, KH, KW, stride, pad, out=None):
    """Tensor-core convolution with an already packed (Cout, KH, KW, C) weight (no bias / activation)."""
    N, H, W, C = x.shape
    Cout = pk.shape[0]
    OH, OW = conv_out(H, KH, stride, pad), conv_out(W, KW, stride, pad)
    if out is None:
        out = torch.empty((N, OH, OW, Cout), device=x.device, dtype=torch.float32)
    call("avl_tc_conv2d_fwd", fptr(x), N, H, W, C, fptr(pk), Cout, KH, KW, stride, pad, None, None, None, 0, 0,
         out.data_ptr(), Cout, stream())
    return out


_wgrad_ws = {}
_tc_backward = [True]


def set_tc_backward(on) -> bool:
    """Tensor-core data / weight gradients of the convolutions (csrc/conv_bwd_tc.cu) on / off (off: fp32 SIMT kernels
    of csrc/nn_bwd.cu).  Returns the previous setting."""
    old = _tc_backward[0]
    _tc_backward[0] = bool(on)
    return old


def zero_upsample2(gy, H, W):
    """(N, OH, OW, C) -> (N, H, W, C) with gy at the even positions and zeros elsewhere (stride-2 data gradients)."""
    N, OH, OW, C = gy.shape
    up = torch.empty((N, H, W, C), device=gy.device, dtype=torch.float32)
    call("avl_zero_upsample2", fptr(gy), fptr(up), N, OH, OW, C, H, W, stream())
    return up


_S2_ROW = ((1, None), (2, 0))  # kernel row read by output parity pa through dy row a + u: _S2_ROW[pa][u]


def _dgrad_s2_weight(w):
    """(Cout, Cin, 3, 3) -> [4 * Cin][4 * Cout] weight of the 2x2-tap data-gradient convolution (row (pa, pb, ci), column
    (u, v, co)); TF32-rounded, cached on the parameter per version like ``_packed_weight``."""
    owner = w._base if w._base is not None else w
    cache = getattr(owner, "_avl_packed", None)
    if cache is None:
        cache = {}
        try:
            owner._avl_packed = cache
        except AttributeError:
            pass
    key = (tuple(w.shape), "s2", w.data_ptr())
    hit = cache.get(key)
    if hit is not None and hit[0] == w._version:
        return hit[1]
    Cout, Cin = w.shape[0], w.shape[1]
    wd = w.detach()
    w2 = torch.zeros((2, 2, Cin, 2, 2, Cout), device=w.device, dtype=torch.float32)
    for pa in range(2):
        for u in range(2):
            r = _S2_ROW[pa][u]
            if r is None:
                continue
            for pb in range(2):
                for v in range(2):
                    sx = _S2_ROW[pb][v]
                    if sx is not None:
                        w2[pa, pb, :, u, v, :] = wd[:, :, r, sx].t()
    w2 = round_to_tf32(w2.view(4 * Cin, 4 * Cout))
    cache[key] = (w._version, w2)
    return w2


def conv2d_dgrad_tc(gy, w, H, W, stride, pad):
    """dx (N, H, W, C) of y = conv(x, w, stride, pad) on the forward tensor-core kernels: a stride-1 convolution of gy
    (zero-upsampled when the forward stride was 2) with the flipped, channel-transposed weight.  Returns None when
    the shape is not covered."""
    Cout, C, KH, KW = w.shape
    if stride not in (1, 2) or Cout % 4 or pad > KH - 1 or pad > KW - 1:
        return None
    cp = (C + 3) // 4 * 4
    if (stride == 2 and KH == 3 and KW == 3 and pad == 1 and H % 2 == 0 and W % 2 == 0 and C % 16 == 0 and Cout >= 16
            and gy.shape[1] * 2 == H and gy.shape[2] * 2 == W and gy.is_contiguous()):
        gx = torch.empty((gy.shape[0], H, W, C), device=gy.device, dtype=torch.float32)
        rc = _lib.lib().avl_tc_conv2d_dgrad_s2(fptr(gy), gy.shape[0], gy.shape[1], gy.shape[2], Cout, fptr(_dgrad_s2_weight(w)), C,
                                               fptr(gx), stream())
        if rc == 0:
            return gx
        if rc != -2:
            _lib.check(rc, "avl_tc_conv2d_dgrad_s2")
    pk = _packed_weight(w, Cout, dgrad=True)  # (C, KH, KW, Cout)
    if cp != C:  # the result is cropped below; pad the packed rows instead of the activations
        pk = torch.nn.functional.pad(pk, (0, 0, 0, 0, 0, 0, 0, cp - C))
    if stride == 1:
        gx = _conv2d_packed(gy, pk, KH, KW, 1, KH - 1 - pad)
    elif KH == 1 and KW == 1:
        gx = zero_upsample2(_conv2d_packed(gy, pk, 1, 1, 1, 0), H, W)
    else:
        up = zero_upsample2(gy, H - KH + 1 + 2 * pad, W - KW + 1 + 2 * pad)
        gx = _conv2d_packed(up, pk, KH, KW, 1, KH - 1 - pad)
    assert gx.shape[1] == H and gx.shape[2] == W, (gx.shape, H, W)
    return gx if cp == C else gx[..., :C].contiguous()


def conv2d_wgrad_tc(x, gy, w_shape, stride, pad):
    """dw (Cout, Cw, KH, KW) on the tensor cores (csrc/conv_bwd_tc.cu); None when the shape is not covered."""
    N, H, W, Cx = x.shape
    Cout, Cw, KH, KW = w_shape
    need = int(_lib.lib().avl_tc_conv2d_wgrad_workspace(N, H, W, Cx, Cout, KH, KW, stride, pad))
    if need < 0:
        return None
    ws = _wgrad_ws.get(x.device)
    if ws is None or ws.numel() < need:
        ws = _wgrad_ws[x.device] = torch.empty(max(need, 1 << 20), device=x.device, dtype=torch.float32)
    gw = torch.empty(w_shape, device=x.device, dtype=torch.float32)
    rc = _lib.lib().avl_tc_conv2d_wgrad(fptr(x), fptr(gy), fptr(gw), N, H, W, Cx, Cw, Cout, KH, KW, stride, pad,
                                        ws.data_ptr(), ws.numel(), stream())
    if rc == -2:
        return None
    _lib.check(rc, "avl_tc_conv2d_wgrad")
    return gw


class _ConvFn(torch.autograd.Function):
    """conv2d (+bias, +ReLU).  Backward: data gradient = a forward tensor-core convolution with the flipped weight,
    weight gradient = the strip-staged TF32 kernel of csrc/conv_bwd_tc.cu; shapes they do not cover (and
    ``set_tensor_cores(0)``) use the fp32 SIMT kernels of csrc/nn_bwd.cu."""

    @staticmethod
    def forward(ctx, x, w, bias, stride, pad, relu):
        x = x.contiguous()
        if tensor_cores_enabled() and x.shape[-1] % 4:  # save the channel-padded input: the weight gradient reads it
            x = pad_channels(x, (x.shape[-1] + 3) // 4 * 4)
        y = _conv2d_raw(x, w, bias, stride, pad, relu)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.cfg = (stride, pad, relu, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        stride, pad, relu, has_bias = ctx.cfg
        N, H, W, C = x.shape
        Cout, Cw, KH, KW = w.shape
        gy = gy.contiguous()
        if relu:
            gy = gy.clone()
            call("avl_relu_mask", fptr(gy), Cout, fptr(y), Cout, gy.numel() // Cout, Cout, stream())
        tc = _tc_backward[0] and tensor_cores_enabled() and gy.data_ptr() % 16 == 0
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            if tc and C == Cw and N * H * W >= _tc_min_rows[0]:
                gx = conv2d_dgrad_tc(gy, w, H, W, stride, pad)
            if gx is None:  # (a channel-padded input gets the gradient of its Cw real channels)
                gx = torch.empty((N, H, W, Cw), device=x.device, dtype=torch.float32)
                call("avl_conv2d_dgrad", fptr(gy), fptr(w.contiguous()), fptr(gx), N, H, W, Cw, Cout, KH, KW, stride, pad,
                     0, stream())
        need_w, need_b = ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]
        if need_w and tc:
            gw = conv2d_wgrad_tc(x, gy, tuple(w.shape), stride, pad)
            if gw is not None and need_b:
                gb = gy.reshape(-1, Cout).sum(0)
                need_b = False
        if (need_w and gw is None) or need_b:
            xs = x if C == Cw else x[..., :Cw].contiguous()
            gw_ = torch.zeros_like(w, memory_format=torch.contiguous_format) if (need_w and gw is None) else None
            gb = torch.zeros(Cout, device=x.device, dtype=torch.float32) if need_b else gb
            call("avl_conv2d_wgrad", fptr(xs), fptr(gy), fptr(gw_), fptr(gb) if need_b else None, N, H, W, Cw, Cout, KH,
                 KW, stride, pad, stream())
            if gw is None:
                gw = gw_
        return gx, gw, gb, None, None, None


class _LinearFlatFn(torch.autograd.Function):
    """nn.Linear over the NCHW-flattened map, computed from the NHWC tensor (forward: a convolution whose kernel covers
    the whole map).  Backward as two dense products on the (O, H, W, C)-ordered weight: dx = dy @ Wp (3xTF32 tcgen05
    GEMM), dWp = dy^T @ x_flat (``avl_tc_wgrad_3x``), permuted back to the reference's (O, C*H*W) order."""

    @staticmethod
    def forward(ctx, x, w, bias, relu):
        x = x.contiguous()
        N, H, W, C = x.shape
        O = w.shape[0]
        y = _conv2d_raw(x, w.view(O, C, H, W), bias, 1, 0, relu).view(N, O)
        ctx.save_for_backward(x, w, y if relu else None)
        ctx.cfg = (relu, bias is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, w, y = ctx.saved_tensors
        relu, has_bias = ctx.cfg
        N, H, W, C = x.shape
        O, K = w.shape[0], H * W * C
        gy = gy.contiguous()
        if relu:
            gy = gy.clone()
            call("avl_relu_mask", fptr(gy), O, fptr(y), O, N, O, stream())
        gx = gw = gb = None
        wp = None
        if ctx.needs_input_grad[0]:
            wp = w.detach().view(O, C, H, W).permute(0, 2, 3, 1).contiguous().view(O, K)  # (O, HWC), exact fp32
            gx = torch.empty((N, K), device=x.device, dtype=torch.float32)
            rc = -2
            if tensor_cores_enabled() and N >= 512:
                rc = _lib.lib().avl_tc_gemm_3x(fptr(gy), O, fptr(wp), K, 1, fptr(gx), K, N, K, O, None, None, 0, 0, None,
                                               stream())
                if rc not in (0, -2):
                    _lib.check(rc, "avl_tc_gemm_3x")
            if rc == -2:
                call("avl_gemm", fptr(gy), O, 1, fptr(wp), 1, K, fptr(gx), K, N, K, O, None, 0, 0, 1, stream())
            gx = gx.view(N, H, W, C)
        if ctx.needs_input_grad[1]:
            gwp = torch.zeros((O, K), device=x.device, dtype=torch.float32)
            xf = x.view(N, K)
            rc = -2
            if tensor_cores_enabled() and N >= 512:
                rc = _lib.lib().avl_tc_wgrad_3x(fptr(gy), O, fptr(xf), K, fptr(gwp), K, N, O, K, None, stream())
                if rc not in (0, -2):
                    _lib.check(rc, "avl_tc_wgrad_3x")
            if rc == -2:
                call("avl_gemm", fptr(gy), 1, O, fptr(xf), 1, K, fptr(gwp), K, O, K, N, None, 0, 0, 1, stream())
            gw = gwp.view(O, H, W, C).permute(0, 3, 1, 2).reshape(O, K)
        if has_bias and ctx.needs_input_grad[2]:
            gb = gy.sum(0)
        return gx, gw, gb, None


class _GroupNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, residual, groups, eps, relu):
        x = x.contiguous()
        y = _groupnorm_raw(x, gamma, beta, groups, eps, relu, residual.contiguous() if residual is not None else None)
        ctx.save_for_backward(x, gamma, y if relu else None)
        ctx.cfg = (groups, eps, relu, residual is not None)
        return y

    @staticmethod
    def backward(ctx, gy):
        x, gamma, y = ctx.saved_tensors
        groups, eps, relu, has_res = ctx.cfg
        N, H, W, C = x.shape
        gy = gy.contiguous()
        gx = torch.empty_like(x) if ctx.needs_input_grad[0] else None
        gres = torch.empty_like(x) if (has_res and ctx.needs_input_grad[3]) else None
        gg = torch.zeros_like(gamma) if ctx.needs_input_grad[1] else None
        gb = torch.zeros_like(gamma) if ctx.needs_input_grad[2] else None
        rc = -2
        if _gn_cluster[0] and C % 4 == 0:
            # one pass over HBM, deterministic parameter gradients (csrc/gn_cluster.cu gn_cluster_bwd_kernel)
            sc = _gn_bwd_scratch.get(x.device)
            if sc is None or sc.numel() < 2 * N * C:
                sc = _gn_bwd_scratch[x.device] = torch.empty(max(2 * N * C, 1 << 16), device=x.device, dtype=torch.float32)
            rc = _lib.lib().avl_groupnorm_bwd_cluster(fptr(x), fptr(y), fptr(gy), fptr(gamma), fptr(gx), fptr(gres),
                                                      fptr(gg), fptr(gb), N, H * W, C, groups, eps, int(relu),
                                                      sc.data_ptr(), stream())
            if rc not in (0, -2):
                _lib.check(rc, "avl_groupnorm_bwd_cluster")
        if rc == -2:
            call("avl_groupnorm_bwd", fptr(x), fptr(y), fptr(gy), fptr(gamma), fptr(gx), fptr(gres), fptr(gg), fptr(gb), N,
                 H * W, C, groups, eps, int(relu), stream())
        return gx, gg, gb, gres, None, None, None


def linear_flat(x_nhwc, w, bias=None, relu=False, out=None):
    """nn.Linear applied to the NCHW-flattened activation, computed from the NHWC tensor: the (O, C*H*W) weight is
    viewed as an (O, C, H, W) kernel covering the whole map (no repacking of the reference's weights)."""
    N, H, W, C = x_nhwc.shape
    O = w.shape[0]
    if _needs_grad(x_nhwc, w, bias):
        return _LinearFlatFn.apply(x_nhwc, w, bias, bool(relu))
    y = _conv2d_raw(x_nhwc, w.view(O, C, H, W), bias, 1, 0, relu, out=out)
    return y.view(N, O) if out is None else out


def linear(x, w, bias=None, relu=False, out=None):
    """y = x @ w.T + b for row-major x (rows, K), w (N, K)."""
    rows, K = x.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty((rows, N), device=x.device, dtype=torch.float32)
    if (rows >= 512 and K % 4 == 0 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0 and w.data_ptr() % 16 == 0
            and tensor_cores_level() >= 2):
        call("avl_tc_gemm", fptr(x), x.stride(0), fptr(w), out.data_ptr(), out.stride(0), rows, N, K, None, fptr(bias),
             None, 0, int(relu), None, stream())
        return out
    call("avl_gemm", fptr(x), x.stride(0), 1, fptr(w), K, 1, out.data_ptr(), out.stride(0), rows, N, K, fptr(bias),
         int(relu), 0, 1, stream())
    return out


def groupnorm(x, gamma, beta, groups=16, eps=1e-5, relu=False, residual=None, out=None):
    """GroupNorm (+ residual) (+ ReLU) on NHWC.  Differentiable when an input requires grad (then ``out`` is
    ignored: the GroupNorm input has to survive for the backward pass)."""
    if _needs_grad(x, gamma, beta, residual):
        return _GroupNormFn.apply(x, gamma, beta, residual, int(groups), float(eps), bool(relu))
    return _groupnorm_raw(x, gamma, beta, groups, eps, relu, residual, out)


def _groupnorm_raw(x, gamma, beta, groups=16, eps=1e-5, relu=False, residual=None, out=None):
    N, H, W, C = x.shape
    if out is None:
        out = torch.empty_like(x)
    if _gn_cluster[0] and C % 4 == 0:
        # one pass over HBM: a thread-block cluster per sample, statistics exchanged through distributed shared memory
        rc = _lib.lib().avl_groupnorm_fwd_cluster(fptr(x), fptr(gamma), fptr(beta), fptr(residual), fptr(out), N, H * W,
                                                  C, groups, float(eps), int(relu), stream())
        if rc == 0:
            return out
        if rc != -2:  # -2: shape outside the cluster kernel -> two-pass kernels below
            _lib.check(rc, "avl_groupnorm_fwd_cluster")
    if C % 4 == 0 and 256 % C == 0 and N * H * W * C >= (1 << 20):
        st = _gn_scratch.get(x.device)
        need = 2 * N * groups + N * C  # doubles: stats + (a, b) float2 per (n, c)
        if st is None or st.numel() < need:
            st = _gn_scratch[x.device] = torch.empty(max(need, 1 << 16), device=x.device, dtype=torch.float64)
        call("avl_groupnorm_fwd_split", fptr(x), fptr(gamma), fptr(beta), fptr(residual), fptr(out), N, H * W, C,
             groups, float(eps), int(relu), st.data_ptr(), stream())
        return out
    call("avl_groupnorm_fwd", fptr(x), fptr(gamma), fptr(beta), fptr(residual), fptr(out), N, H * W, C, groups,
         float(eps), int(relu), stream())
    return out


class IndexedObservation:
    """Rows of a time-major observation store ``(T + 1, N, H, W, C)`` addressed by a flat sample index, WITHOUT
    materialising them: what ``RolloutStorage.recurrent_generator`` hands the policy for the image
    sensors instead of the reference's stacked ``(T * N_mb, H, W, C)`` copies (rollout_storage.py:716-760).  The
    encoders' first kernel (``resize_half``) gathers through the index."""

    def __init__(self, storage, index):
        self.storage = storage.reshape(-1, *storage.shape[2:])  # a view: the store is contiguous
        self.index = index.to(torch.int64).contiguous()

    @property
    def shape(self):
        return (self.index.shape[0],) + tuple(self.storage.shape[1:])

    @property
    def dtype(self):
        return self.storage.dtype

    @property
    def device(self):
        return self.storage.device

    is_cuda = True

    def contiguous(self):
        return self

    def data_ptr(self):
        return self.index.data_ptr()

    def materialize(self):
        return self.storage.index_select(0, self.index)


_RESIZE_DTYPES = {torch.float32: 0, torch.float16: 1, torch.uint8: 2}


def resize_half(x, scale=1.0, c_out=None):
    """Exact 2x2 area mean (x * scale first); ``c_out`` > C appends zero channels (tensor-core conv loader).  ``x``:
    fp32 / fp16 / uint8 NHWC tensor, or an ``IndexedObservation`` (rows gathered through its sample index)."""
    index = None
    if isinstance(x, IndexedObservation):
        index, x = x.index, x.storage
    N, H, W, C = x.shape
    rows = N if index is None else index.shape[0]
    c_out = C if c_out is None else c_out
    y = torch.empty((rows, H // 2, W // 2, c_out), device=x.device, dtype=torch.float32)
    if index is None and x.dtype == torch.float32:
        call("avl_resize_half", fptr(x), fptr(y), N, H, W, C, c_out, float(scale), stream())
        return y
    code = _RESIZE_DTYPES.get(x.dtype)
    if code is None:
        raise _lib.AvlenError(f"resize_half: unsupported observation dtype {x.dtype}")
    call("avl_resize_half_typed", dptr(x), code, dptr(index, torch.int64), fptr(y), rows, H, W, C, c_out, float(scale),
         stream())
    return y


def concat_rgbd(rgb, depth, rgb_scale=1.0 / 255.0):
    N, H, W, _ = (rgb if rgb is not None else depth).shape
    cr = rgb.shape[-1] if rgb is not None else 0
    cd = depth.shape[-1] if depth is not None else 0
    y = torch.empty((N, H, W, cr + cd), device=(rgb if rgb is not None else depth).device, dtype=torch.float32)
    call("avl_concat_rgbd", fptr(rgb), fptr(depth), fptr(y), N * H * W, cr, cd, float(rgb_scale), stream())
    return y


def append_planes(x, extra):
    N, H, W, C = x.shape
    E = extra.shape[1]
    y = torch.empty((N, H, W, C + E), device=x.device, dtype=torch.float32)
    call("avl_append_planes", fptr(x), fptr(extra.contiguous()), fptr(y), N, H * W, C, E, stream())
    return y


def maxpool3x3s2(x):
    N, H, W, C = x.shape
    y = torch.empty((N, conv_out(H, 3, 2, 1), conv_out(W, 3, 2, 1), C), device=x.device, dtype=torch.float32)
    call("avl_maxpool3x3s2", fptr(x), fptr(y), N, H, W, C, stream())
    return y


def avgpool_global(x):
    N, H, W, C = x.shape
    y = torch.empty((N, C), device=x.device, dtype=torch.float32)
    call("avl_avgpool_global", fptr(x), fptr(y), N, H * W, C, stream())
    return y


def onehot_linear(actions, w, bias, out):
    """out[:, :] = one_hot(actions) @ w.T + bias written into a (B, out_dim) column slice."""
    B = actions.shape[0]
    call("avl_onehot_linear", dptr(actions.reshape(B).contiguous(), torch.int64), fptr(w), fptr(bias), out.data_ptr(),
         out.stride(0), B, w.shape[0], w.shape[1], stream())
    return out


def copy_cols(src, dst):
    rows, cols = src.shape
    call("avl_copy_cols", src.data_ptr(), src.stride(0), dst.data_ptr(), dst.stride(0), rows, cols, stream())
    return dst


# ------------------------------------------------------------------------------------------------------------------
# Whole-network ResNet-18 inference behind one C-ABI call (csrc/resnet_fwd.cu)
class ResNetPlan:
    """Host-side plan of one network for ``avl_resnet18_forward``: the 12-int configuration, the device-pointer table
    (rebuilt when a parameter changes or the tensor-core mode flips) and a private activation workspace."""

    def __init__(self, cfg, tensors_fn, module=None):
        import ctypes
        self._ct = ctypes
        self.cfg = (ctypes.c_int * 12)(*cfg)
        self._tensors_fn = tensors_fn  # (use_tc) -> list of 77 tensors / None, plus a version key
        self._key = None
        self._keep = None
        self._table = None
        self._ws = None
        # fast path: (owner dict, name) of every parameter / buffer of ``module``, collected once.  Re-walking the
        # module tree (named_parameters) on every call cost ~0.5 ms of host time per network and rollout step.
        self._refs = None
        self._module = module
        self._fp = None
        self._cast = None

    def _fingerprint(self, use_tc):
        if self._refs is None:
            refs = []
            for m in self._module.modules():
                refs += [(m._parameters, k) for k in m._parameters] + [(m._buffers, k) for k in m._buffers]
            self._refs = refs
        ts = [d[k] for d, k in self._refs]
        ts = [t for t in ts if t is not None]
        return (use_tc, [id(t) for t in ts], [t._version for t in ts], [t.data_ptr() for t in ts])

    def any_requires_grad(self):
        if self._refs is None:
            self._fingerprint(0)
        for d, k in self._refs:
            t = d[k]
            if t is not None and t.requires_grad:
                return True
        return False

    def table(self, use_tc):
        if self._module is not None:
            fp = self._fingerprint(use_tc)
            if fp == self._fp:
                return self._cast
            self._fp = None
        tensors, key = self._tensors_fn(use_tc)
        key = (use_tc, key, tuple(0 if t is None else t.data_ptr() for t in tensors))
        if key != self._key:
            for t in tensors:
                if t is not None and (t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous()):
                    raise _lib.AvlenError("ResNet parameters must be contiguous fp32 CUDA tensors")
            self._table = (self._ct.c_void_p * len(tensors))(*[None if t is None else t.data_ptr() for t in tensors])
            self._keep, self._key = tensors, key
        self._cast = self._ct.cast(self._table, self._ct.c_void_p)
        if self._module is not None:
            self._fp = self._fingerprint(use_tc)
        return self._cast

    def workspace(self, N, H, W, device):
        nbytes = int(_lib.lib().avl_resnet18_workspace_bytes(N, H, W, self._ct.cast(self.cfg, self._ct.c_void_p)))
        if nbytes < 0:
            raise _lib.AvlenError("invalid ResNet configuration")
        if self._ws is None or self._ws.numel() < nbytes or self._ws.device != device:
            self._ws = torch.empty(nbytes, dtype=torch.uint8, device=device)
        return self._ws

    def cfg_ptr(self):
        return self._ct.cast(self.cfg, self._ct.c_void_p)


def _prep_net_input(x, use_tc):
    x = x.contiguous()
    if use_tc and x.shape[-1] % 4:
        x = pad_channels(x, (x.shape[-1] + 3) // 4 * 4)
    return x


def resnet18_forward(plan, x, out, eps=1e-5):
    """One network: x (N, H, W, C) NHWC -> out (N, out_dim) (a column slice is fine)."""
    use_tc = int(tensor_cores_enabled())
    x = _prep_net_input(x, use_tc)
    N, H, W, C = x.shape
    call("avl_resnet18_forward", fptr(x), N, H, W, C, plan.cfg_ptr(), float(eps), plan.table(use_tc), out.data_ptr(),
         out.stride(0), use_tc, plan.workspace(N, H, W, x.device).data_ptr(), stream())
    return out


def resnet18_forward_pair(plan0, x0, out0, plan1, x1, out1, eps=1e-5):
    """Two independent networks on two streams (fork / join around the caller's stream)."""
    use_tc = int(tensor_cores_enabled())
    x0, x1 = _prep_net_input(x0, use_tc), _prep_net_input(x1, use_tc)
    N, H, W, C0 = x0.shape
    assert x1.shape[:3] == x0.shape[:3]
    call("avl_resnet18_forward_pair", fptr(x0), fptr(x1), N, H, W, C0, x1.shape[3], plan0.cfg_ptr(), plan1.cfg_ptr(),
         float(eps), plan0.table(use_tc), plan1.table(use_tc), out0.data_ptr(), out1.data_ptr(), out0.stride(0),
         out1.stride(0), use_tc, plan0.workspace(N, H, W, x0.device).data_ptr(),
         plan1.workspace(N, H, W, x0.device).data_ptr(), stream())
    return out0, out1

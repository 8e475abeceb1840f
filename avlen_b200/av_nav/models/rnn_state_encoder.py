"""RNNStateEncoder (ss_baselines/av_nav/models/rnn_state_encoder.py:15-149): GRU(input -> hidden, 1 layer) whose
state is multiplied by the not-done mask before every step.

``self.rnn`` is an ``nn.GRU`` kept as the parameter container (names ``rnn.weight_ih_l0`` ... as in the reference's
``state_dict``); the arithmetic runs on the gate-fused CUDA kernels of csrc/nn_bwd.cu: ONE input-projection GEMM for
all T*N rows, then per step one recurrent GEMM + one fused gate kernel.  The reference's ``seq_forward`` splits the
sequence at steps where any env was reset (a ``nonzero().cpu()`` host sync, :111-120) and calls cuDNN per chunk;
multiplying the state by ``mask_t`` at every step is the same function with no host round trip.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from ... import _lib
from ... import nn as K


class _GruFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, h0, masks, w_ih, w_hh, b_ih, b_hh, T, need_grad):
        TN, I = x.shape
        N = TN // T
        H = h0.shape[1]
        nbytes = int(_lib.lib().avl_gru_workspace_bytes(T, N, I, H, int(need_grad)))
        ws = torch.empty(nbytes, dtype=torch.uint8, device=x.device)
        out = torch.empty((TN, H), device=x.device, dtype=torch.float32)
        h_last = torch.empty((N, H), device=x.device, dtype=torch.float32)
        _lib.call("avl_gru_forward", T, N, I, H, _lib.fptr(x), _lib.fptr(h0), _lib.fptr(masks), _lib.fptr(w_ih),
                  _lib.fptr(w_hh), _lib.fptr(b_ih), _lib.fptr(b_hh), _lib.fptr(out), _lib.fptr(h_last),
                  ws.data_ptr(), int(need_grad), _lib.stream())
        if need_grad:
            ctx.save_for_backward(x, masks, w_ih, w_hh)
            ctx.ws, ctx.dims = ws, (T, N, I, H)
        return out, h_last

    @staticmethod
    def backward(ctx, g_out, g_hlast):
        x, masks, w_ih, w_hh = ctx.saved_tensors
        T, N, I, H = ctx.dims
        dev = x.device
        ni = ctx.needs_input_grad
        dx = torch.empty_like(x) if ni[0] else None
        dh0 = torch.empty((N, H), device=dev, dtype=torch.float32) if ni[1] else None
        dw_ih = torch.zeros_like(w_ih) if ni[3] else None
        dw_hh = torch.zeros_like(w_hh) if ni[4] else None
        db_ih = torch.zeros(3 * H, device=dev, dtype=torch.float32) if ni[5] else None
        db_hh = torch.zeros(3 * H, device=dev, dtype=torch.float32) if ni[6] else None
        g_out = g_out.contiguous() if g_out is not None else None
        g_hlast = g_hlast.contiguous() if g_hlast is not None else None
        _lib.call("avl_gru_backward", T, N, I, H, _lib.fptr(x), _lib.fptr(masks), _lib.fptr(w_ih), _lib.fptr(w_hh),
                  _lib.fptr(g_out), _lib.fptr(g_hlast), _lib.fptr(dx), _lib.fptr(dh0), _lib.fptr(dw_ih),
                  _lib.fptr(dw_hh), _lib.fptr(db_ih), _lib.fptr(db_hh), ctx.ws.data_ptr(), _lib.stream())
        return dx, dh0, None, dw_ih, dw_hh, db_ih, db_hh, None, None


class RNNStateEncoder(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, num_layers: int = 1, rnn_type: str = "GRU"):
        super().__init__()
        if rnn_type != "GRU" or num_layers != 1:
            raise _lib.AvlenError("the CUDA state encoder is built for a 1-layer GRU (every av_nav yaml)")
        self._num_recurrent_layers = num_layers
        self._rnn_type = rnn_type
        self._hidden_size = hidden_size
        self.rnn = nn.GRU(input_size=input_size, hidden_size=hidden_size, num_layers=num_layers)
        self.layer_init()

    def layer_init(self):
        for name, param in self.rnn.named_parameters():
            if "weight" in name:
                nn.init.orthogonal_(param)
            elif "bias" in name:
                nn.init.constant_(param, 0)

    @property
    def num_recurrent_layers(self):
        return self._num_recurrent_layers

    def _run(self, x, hidden_states, masks, T):
        r = self.rnn
        need = torch.is_grad_enabled() and (x.requires_grad or hidden_states.requires_grad or
                                            any(p.requires_grad for p in r.parameters()))
        out, h = _GruFn.apply(x.contiguous(), hidden_states[0].contiguous(), masks.reshape(-1).float().contiguous(),
                              r.weight_ih_l0, r.weight_hh_l0, r.bias_ih_l0, r.bias_hh_l0, T, need)
        return out, h.unsqueeze(0)

    def single_forward(self, x, hidden_states, masks):
        """:80-90 — x (N, I), hidden_states (1, N, H), masks (N, 1)."""
        return self._run(x, hidden_states, masks, 1)

    def seq_forward(self, x, hidden_states, masks):
        """:92-143 — x (T*N, I) time-major, hidden_states (1, N, H), masks (T*N, 1)."""
        n = hidden_states.size(1)
        return self._run(x, hidden_states, masks, x.size(0) // n)

    def forward(self, x, hidden_states, masks):
        if x.size(0) == hidden_states.size(1):
            return self.single_forward(x, hidden_states, masks)
        return self.seq_forward(x, hidden_states, masks)

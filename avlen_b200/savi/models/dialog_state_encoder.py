"""DialogStateEncoder (ss_baselines/savi/models/dialog_state_encoder.py:18-160) on hand-written CUDA kernels.

Same constructor, parameter / buffer names (``fusion_encoder.{0,2}``, ``dialog_transformer.*`` in the
``torch.nn.Transformer`` layout, ``pos_encode.pe``) and call signature as the reference.  Forward and backward are
single C-ABI calls (``avl_dialog_forward`` / ``avl_dialog_backward``, csrc/smt.cu): the valid slots of the K-slot
state memory plus the current SMT output are packed into token rows on the device, fused with the dialog embedding,
offset by the sinusoid row of ``agent_step`` and run through the shared transformer block.
"""
from __future__ import annotations

import ctypes
import math

import torch
import torch.nn as nn

from ... import _lib
from .smt_state_encoder import SMT_PARAM_KEYS, IndexedMemory, _PtrTable

# transformer entries in the order of csrc/smt.cu's TP_* enum, then the fusion MLP (DP_*)
DIALOG_PARAM_KEYS = [k.replace("transformer.", "dialog_transformer.", 1) for k in SMT_PARAM_KEYS
                     if k.startswith("transformer.")] + \
                    ["fusion_encoder.0.weight", "fusion_encoder.0.bias", "fusion_encoder.2.weight", "fusion_encoder.2.bias"]

_c, _p = ctypes.c_int, ctypes.c_void_p
_lib.register({
    "avl_dialog_param_count": [],
    "avl_dialog_workspace_bytes": [_c, _c, _c, _c],
    "avl_dialog_forward": [_c, _c, _c, _p, _p, _c, _p, _p, _p, _p, _p, _c, _p, _p, _p, _p, _c, _p],
    "avl_dialog_backward": [_c, _c, _c, _c, _p, _p, _p, _p, _p, _p, _p, _p, _p],
}, {"avl_dialog_workspace_bytes": ctypes.c_longlong})


class PositionalEncoding(nn.Module):
    """dialog_state_encoder.py:18-40: sinusoid table indexed by the per-sample agent step (the add happens inside
    the kernel; this module only owns the ``pe`` buffer so ``state_dict`` keys match)."""

    def __init__(self, d_model: int, dropout: float = 0.1, max_len: int = 5000):
        super().__init__()
        if dropout != 0.0:
            raise _lib.AvlenError("PositionalEncoding dropout must be 0.0 (dialog_state_encoder.py:100)")
        position = torch.arange(max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-math.log(10000.0) / d_model))
        pe = torch.zeros(max_len, 1, d_model)
        pe[:, 0, 0::2] = torch.sin(position * div_term)
        pe[:, 0, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe)


class _DialogFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x, memory, env_index, masks, d_emb, agent_step, goal, need_grad, *params):
        B, D = x.shape
        K = memory.shape[0] if memory is not None else 0
        n_mem = memory.shape[1] if memory is not None else 0
        ws = enc._workspace(B, K, D, need_grad)
        out = torch.empty((B, D), device=x.device, dtype=torch.float32)
        pe = enc.pos_encode.pe
        _lib.call("avl_dialog_forward", B, K, D, _lib.fptr(x), _lib.fptr(memory), n_mem,
                  _lib.dptr(env_index, torch.int32), _lib.fptr(masks), _lib.fptr(d_emb),
                  _lib.dptr(agent_step, torch.int32), _lib.fptr(pe), pe.shape[0], _lib.fptr(goal),
                  ctypes.cast(enc._ptab.get(params), _p), _lib.fptr(out), ws.data_ptr(), int(need_grad), _lib.stream())
        if need_grad:
            # single training workspace per module: refuse a backward whose activations were overwritten
            enc._train_generation = ctx.generation = getattr(enc, "_train_generation", 0) + 1
            ctx.enc, ctx.dims, ctx.ws, ctx.params, ctx.goal = enc, (B, K, D), ws, params, goal
            ctx.has_dialog = d_emb is not None
            ctx.need = (x.requires_grad, d_emb is not None and d_emb.requires_grad, goal.requires_grad)
        return out

    @staticmethod
    def backward(ctx, gout):
        enc = ctx.enc
        if enc._train_generation != ctx.generation:
            raise _lib.AvlenError("DialogStateEncoder: another grad-enabled forward of this module ran before this "
                                  "backward (the saved activations share one workspace); run backward per forward")
        B, K, D = ctx.dims
        params = ctx.params
        grads = []
        for p in params:
            if p.requires_grad:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                grads.append(p.grad)
            else:
                grads.append(None)
        need_x, need_d, need_g = ctx.need
        mk = lambda need: torch.empty((B, D), device=gout.device, dtype=torch.float32) if need else None
        dx, dd, dgoal = mk(need_x), mk(need_d), mk(need_g)
        _lib.call("avl_dialog_backward", B, K, D, int(ctx.has_dialog), _lib.fptr(ctx.goal),
                  ctypes.cast(enc._ptab.get(params), _p), ctypes.cast(enc._gtab.get(grads), _p),
                  _lib.fptr(gout.contiguous()), _lib.fptr(dx), _lib.fptr(dd), _lib.fptr(dgoal), ctx.ws.data_ptr(),
                  _lib.stream())
        return (None, dx, None, None, None, dd, None, dgoal, None) + (None,) * len(params)


class DialogStateEncoder(nn.Module):
    def __init__(self, input_size: int, nhead: int = 8, num_encoder_layers: int = 1, num_decoder_layers: int = 1,
                 dim_feedforward: int = 256, dropout: float = 0.1, activation: str = "relu",
                 pretraining: bool = False, **_unused):
        super().__init__()
        if (nhead, num_encoder_layers, num_decoder_layers, dim_feedforward, activation) != (8, 1, 1, 256, "relu"):
            raise _lib.AvlenError("the CUDA dialog encoder is built for nhead=8, 1+1 layers, hidden 256, relu")
        if dropout != 0.0:
            raise _lib.AvlenError("dropout must be 0.0 (as in every SMT yaml of the reference)")
        if input_size != 2 * dim_feedforward:
            raise _lib.AvlenError("input_size must be hidden + hidden (policy.py:767)")
        self._input_size, self._nhead = input_size, nhead
        self._num_encoder_layers, self._num_decoder_layers = num_encoder_layers, num_decoder_layers
        self._dim_feedforward, self._dropout, self._activation = dim_feedforward, dropout, activation
        self._pretraining = pretraining
        self.fusion_encoder = nn.Sequential(nn.Linear(input_size, dim_feedforward), nn.ReLU(),
                                            nn.Linear(dim_feedforward, dim_feedforward))
        # parameter container only (names / shapes / initialisation of torch.nn.Transformer); never called
        self.dialog_transformer = nn.Transformer(d_model=dim_feedforward, nhead=nhead, num_encoder_layers=1,
                                                 num_decoder_layers=1, dim_feedforward=dim_feedforward,
                                                 dropout=dropout, activation=activation)
        self.pos_encode = PositionalEncoding(d_model=dim_feedforward, dropout=0.0, max_len=100)
        self._ptab, self._gtab = _PtrTable(), _PtrTable()
        self._ws = {}

    @property
    def hidden_state_size(self):
        return self._dim_feedforward

    def _workspace(self, B, K, D, bwd):
        nbytes = int(_lib.lib().avl_dialog_workspace_bytes(B, K, D, int(bwd)))
        key = "train" if bwd else "infer"
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = self._ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=self.pos_encode.pe.device)
        return ws

    def _params(self):
        sd = dict(self.named_parameters())
        return [sd[k] for k in DIALOG_PARAM_KEYS]

    def single_forward(self, x, memory_state, memory_masks, d_emb, agent_step, goal=None):
        if goal is None:
            raise _lib.AvlenError("goal is required (dialog_state_encoder.py:146 asserts it)")
        env_index = None
        if isinstance(memory_state, IndexedMemory):
            env_index, memory_state = memory_state.env_index, memory_state.memory
        assert x.size(0) == (env_index.shape[0] if env_index is not None else memory_state.size(1))
        if memory_masks.shape[1] != memory_state.shape[0]:
            raise _lib.AvlenError("memory_masks must have one column per state-memory slot")
        need_grad = torch.is_grad_enabled() and (
            x.requires_grad or goal.requires_grad or (d_emb is not None and d_emb.requires_grad)
            or any(p.requires_grad for p in self.parameters()))
        return _DialogFunction.apply(self, x.contiguous(), memory_state.contiguous(), env_index,
                                     memory_masks.contiguous(), None if d_emb is None else d_emb.contiguous(),
                                     agent_step.reshape(-1).to(torch.int32).contiguous(), goal.contiguous(), need_grad,
                                     *self._params())

    def forward(self, x, memory_state, *args, **kwargs):
        return self.single_forward(x, memory_state, *args, **kwargs)

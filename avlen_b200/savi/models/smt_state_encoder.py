"""SMTStateEncoder (ss_baselines/savi/models/smt_state_encoder.py:23-280) on hand-written CUDA kernels.

Same constructor, parameter names/shapes (``pose_encoder``, ``fusion_encoder.{0,2}``, ``transformer.*`` with the
``torch.nn.Transformer`` layout) and call signature as the reference, so checkpoints load unchanged.  The forward
and backward passes are single C-ABI calls (``avl_smt_forward`` / ``avl_smt_backward``) wrapped in one
``torch.autograd.Function``.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Tuple

import torch
import torch.nn as nn

from ... import _lib

# order of the device-pointer table expected by csrc/smt.cu (enum TP_* then SP_*)
SMT_PARAM_KEYS = [
    "transformer.encoder.layers.0.self_attn.in_proj_weight", "transformer.encoder.layers.0.self_attn.in_proj_bias",
    "transformer.encoder.layers.0.self_attn.out_proj.weight", "transformer.encoder.layers.0.self_attn.out_proj.bias",
    "transformer.encoder.layers.0.linear1.weight", "transformer.encoder.layers.0.linear1.bias",
    "transformer.encoder.layers.0.linear2.weight", "transformer.encoder.layers.0.linear2.bias",
    "transformer.encoder.layers.0.norm1.weight", "transformer.encoder.layers.0.norm1.bias",
    "transformer.encoder.layers.0.norm2.weight", "transformer.encoder.layers.0.norm2.bias",
    "transformer.encoder.norm.weight", "transformer.encoder.norm.bias",
    "transformer.decoder.layers.0.self_attn.in_proj_weight", "transformer.decoder.layers.0.self_attn.in_proj_bias",
    "transformer.decoder.layers.0.self_attn.out_proj.weight", "transformer.decoder.layers.0.self_attn.out_proj.bias",
    "transformer.decoder.layers.0.multihead_attn.in_proj_weight",
    "transformer.decoder.layers.0.multihead_attn.in_proj_bias",
    "transformer.decoder.layers.0.multihead_attn.out_proj.weight",
    "transformer.decoder.layers.0.multihead_attn.out_proj.bias",
    "transformer.decoder.layers.0.linear1.weight", "transformer.decoder.layers.0.linear1.bias",
    "transformer.decoder.layers.0.linear2.weight", "transformer.decoder.layers.0.linear2.bias",
    "transformer.decoder.layers.0.norm1.weight", "transformer.decoder.layers.0.norm1.bias",
    "transformer.decoder.layers.0.norm2.weight", "transformer.decoder.layers.0.norm2.bias",
    "transformer.decoder.layers.0.norm3.weight", "transformer.decoder.layers.0.norm3.bias",
    "transformer.decoder.norm.weight", "transformer.decoder.norm.bias",
    "pose_encoder.weight", "pose_encoder.bias",
    "fusion_encoder.0.weight", "fusion_encoder.0.bias", "fusion_encoder.2.weight", "fusion_encoder.2.bias",
]

_lib.register({
    "avl_smt_param_count": [],
    "avl_smt_workspace_bytes": [ctypes.c_int] * 6,
    "avl_smt_forward": [ctypes.c_int] * 7 + [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                            ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_void_p],
    "avl_smt_backward": [ctypes.c_int] * 6 + [ctypes.c_void_p] * 8,
    "avl_smt_status": [ctypes.c_int] * 4 + [ctypes.c_void_p, ctypes.POINTER(ctypes.c_int),
                                           ctypes.POINTER(ctypes.c_int)],
}, {"avl_smt_workspace_bytes": ctypes.c_longlong})


class IndexedMemory:
    """A (M, B, dim) external-memory batch expressed WITHOUT materialising it: row b of the batch reads
    ``memory[:, env_index[b]]`` of the single-copy ring buffer (M, N, dim).  Produced by
    ``RolloutStorage.recurrent_generator`` instead of the reference's stacked copies (rollout_storage.py:683-772)."""

    def __init__(self, memory: torch.Tensor, env_index: torch.Tensor):
        self.memory = memory
        self.env_index = env_index.to(torch.int32).contiguous()

    def size(self, d):
        return (self.memory.shape[0], self.env_index.shape[0], self.memory.shape[2])[d]

    @property
    def shape(self):
        return (self.memory.shape[0], self.env_index.shape[0], self.memory.shape[2])

    def materialize(self):
        return self.memory[:, self.env_index.long()]


class _PtrTable:
    """Host array of device pointers handed to the C-ABI (rebuilt only when a parameter moves)."""

    def __init__(self):
        self.key = None
        self.arr = None

    def get(self, tensors):
        key = tuple(0 if t is None else t.data_ptr() for t in tensors)
        if key != self.key:
            self.arr = (ctypes.c_void_p * len(key))(*[k or None for k in key])
            self.key = key
        return self.arr


class _SMTFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, enc, x, memory, env_index, masks, goal, need_grad, *params):
        B, F = x.shape
        M = memory.shape[0] if memory is not None else 0
        n_mem = memory.shape[1] if memory is not None else 0
        D = enc._dim_feedforward
        pi = enc._pose_indices[0]
        rows_cap = enc._rows_cap(B, M)
        need_dx = bool(need_grad and x.requires_grad)
        ws = enc._workspace(B, rows_cap, F, D, need_grad, need_dx)
        out = torch.empty((B, D), device=x.device, dtype=torch.float32)
        ptab = enc._ptab.get(params)
        _lib.call("avl_smt_forward", B, M, F, D, pi, int(enc._pretraining), rows_cap, _lib.fptr(x),
                  _lib.fptr(memory), n_mem, _lib.dptr(env_index, torch.int32), _lib.fptr(masks), _lib.fptr(goal),
                  ctypes.cast(ptab, ctypes.c_void_p), _lib.fptr(out), ws.data_ptr(), int(need_grad), int(need_dx),
                  _lib.stream())
        if need_grad:
            # the activations live in the module's single training workspace: a second grad-enabled forward before
            # this one's backward would overwrite them -> stamp a generation and refuse a stale backward
            enc._train_generation = ctx.generation = getattr(enc, "_train_generation", 0) + 1
            ctx.enc, ctx.dims, ctx.ws, ctx.params, ctx.goal = enc, (B, M, F, D, pi, rows_cap), ws, params, goal
            ctx.need_dx, ctx.need_dgoal = need_dx, goal.requires_grad
        return out

    @staticmethod
    def backward(ctx, gout):
        enc = ctx.enc
        if enc._train_generation != ctx.generation:
            raise _lib.AvlenError("SMTStateEncoder: another grad-enabled forward of this module ran before this "
                                  "backward (the saved activations share one workspace); run backward per forward")
        B, M, F, D, pi, rows_cap = ctx.dims
        params = ctx.params
        grads = []
        for p in params:
            if p.requires_grad:
                if p.grad is None:
                    p.grad = torch.zeros_like(p)
                grads.append(p.grad)
            else:
                grads.append(None)
        gtab = enc._gtab.get(grads)
        dx = torch.zeros((B, F), device=gout.device, dtype=torch.float32) if ctx.need_dx else None
        dgoal = torch.empty((B, D), device=gout.device, dtype=torch.float32) if ctx.need_dgoal else None
        _lib.call("avl_smt_backward", B, M, F, D, pi, rows_cap, _lib.fptr(ctx.goal),
                  ctypes.cast(enc._ptab.get(params), ctypes.c_void_p), ctypes.cast(gtab, ctypes.c_void_p),
                  _lib.fptr(gout.contiguous()), _lib.fptr(dx), _lib.fptr(dgoal), ctx.ws.data_ptr(), _lib.stream())
        # parameter gradients were accumulated in place into p.grad by the kernels
        return (None, dx, None, None, None, dgoal, None) + (None,) * len(params)


class SMTStateEncoder(nn.Module):
    def __init__(self, input_size: int, nhead: int = 8, num_encoder_layers: int = 1, num_decoder_layers: int = 1,
                 dim_feedforward: int = 256, dropout: float = 0.1, activation: str = "relu",
                 pose_indices: Optional[Tuple[int, int]] = None, pretraining: bool = False,
                 query_count_emb_size=32, use_query_count=False):
        super().__init__()
        if (nhead, num_encoder_layers, num_decoder_layers, dim_feedforward, activation) != (8, 1, 1, 256, "relu"):
            raise _lib.AvlenError("the CUDA SMT path is built for nhead=8, 1+1 layers, hidden 256, relu "
                                  "(every SMT yaml of the reference)")
        if dropout != 0.0:
            raise _lib.AvlenError("dropout must be 0.0 (as in every SMT yaml of the reference)")
        if pose_indices is None or pose_indices[1] - pose_indices[0] != 4:
            raise _lib.AvlenError("pose_indices with 4 pose dims are required")
        self._input_size = input_size
        self._nhead, self._num_encoder_layers, self._num_decoder_layers = nhead, num_encoder_layers, num_decoder_layers
        self._dim_feedforward, self._dropout, self._activation = dim_feedforward, dropout, activation
        self._pose_indices = tuple(pose_indices)
        self._pretraining = pretraining
        self._use_pose_encoding = True
        self.pose_encoder = nn.Linear(5, 16)
        fin = input_size + 12
        self.fusion_encoder = nn.Sequential(nn.Linear(fin, dim_feedforward), nn.ReLU(),
                                            nn.Linear(dim_feedforward, dim_feedforward))
        # parameter container only (names/shapes/initialisation of torch.nn.Transformer); never called
        self.transformer = nn.Transformer(d_model=dim_feedforward, nhead=nhead, num_encoder_layers=1,
                                          num_decoder_layers=1, dim_feedforward=dim_feedforward, dropout=dropout,
                                          activation=activation)
        self._ptab, self._gtab = _PtrTable(), _PtrTable()
        self._ws = {}
        self.rows_per_sample_cap: Optional[int] = None  # e.g. capacity + 1; default M + 1

    @property
    def hidden_state_size(self):
        return self._dim_feedforward

    @property
    def pose_indices(self):
        return self._pose_indices

    def _rows_cap(self, B, M):
        per = M + 1 if self.rows_per_sample_cap is None else min(M + 1, self.rows_per_sample_cap)
        return B * (1 if self._pretraining else per)

    def _workspace(self, B, rows_cap, F, D, bwd, need_dx):
        nbytes = int(_lib.lib().avl_smt_workspace_bytes(B, rows_cap, F, D, int(bwd), int(need_dx)))
        key = "train" if bwd else "infer"
        ws = self._ws.get(key)
        if ws is None or ws.numel() < nbytes:
            ws = self._ws[key] = torch.empty(nbytes, dtype=torch.uint8, device=self.pose_encoder.weight.device)
        return ws

    def _params(self):
        # (owner dict, name) of every parameter, resolved once: walking named_parameters() on every call cost ~0.1 ms
        # of host time per rollout step; parameters replaced in their module (load / .to()) are still picked up
        refs = self.__dict__.get("_param_refs")
        if refs is None:
            mods = dict(self.named_modules())
            refs = []
            for k in SMT_PARAM_KEYS:
                owner, _, leaf = k.rpartition(".")
                refs.append((mods[owner]._parameters, leaf))
            self.__dict__["_param_refs"] = refs
        return [d[k] for d, k in refs]

    def last_token_count(self, B, M, F):
        """(synchronising) number of packed token rows and overflow flag of the last inference forward."""
        t, o = ctypes.c_int(0), ctypes.c_int(0)
        key = "train" if "train" in self._ws else "infer"
        _lib.check(_lib.lib().avl_smt_status(B, self._rows_cap(B, M), F, self._dim_feedforward,
                                             self._ws[key].data_ptr(), ctypes.byref(t), ctypes.byref(o)))
        return t.value, o.value

    def single_forward(self, x, memory, memory_masks, goal=None):
        if goal is None:
            raise _lib.AvlenError("goal=None (decoding from memory[-1:]) is not used by any reference policy")
        env_index = None
        if isinstance(memory, IndexedMemory):
            env_index, memory = memory.env_index, memory.memory
        assert x.size(0) == (env_index.shape[0] if env_index is not None else memory.size(1))
        need_grad = torch.is_grad_enabled() and (x.requires_grad or goal.requires_grad or
                                                 any(p.requires_grad for p in self.parameters()))
        return _SMTFunction.apply(self, x.contiguous(), memory.contiguous(), env_index, memory_masks.contiguous(),
                                  goal.contiguous(), need_grad, *self._params())

    def forward(self, x, memory, *args, **kwargs):
        return self.single_forward(x, memory, *args, **kwargs)

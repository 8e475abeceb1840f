"""CLIP ViT-B/32 text tower as the AVLEN dialog policy uses it (ss_baselines/savi/ppo/policy.py:761-762 ``clip.load``,
:844-851 ``self.clip.encode_text(all_dialog).float()``; frozen, ddppo_trainer.py:401-403).

Parameter names / shapes are openai/CLIP's (``token_embedding.weight``, ``positional_embedding``,
``transformer.resblocks.{i}.{ln_1,attn.in_proj_*,attn.out_proj,ln_2,mlp.c_fc,mlp.c_proj}``, ``ln_final``,
``text_projection``) so ``net.clip.*`` entries of a reference checkpoint load by name; the image tower
(``net.clip.visual.*``) is never executed on this path and is not instantiated (load with ``strict=False``).
``encode_text`` is one C-ABI call (``avl_clip_text_forward``): device-side de-duplication of the all-zero dialog
rows, tcgen05 TF32 GEMMs, fused causal attention."""
from __future__ import annotations

import ctypes
from collections import OrderedDict

import torch
import torch.nn as nn

from ... import _lib

_lib.register({
    "avl_clip_text_param_count": [ctypes.c_int],
    "avl_clip_text_workspace_bytes": [ctypes.c_int, ctypes.c_int],
    "avl_clip_text_forward": [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                              ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p],
    "avl_clip_text_forward_f16": [ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p,
                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int,
                                  ctypes.c_void_p],
    "avl_clip_text_status": [ctypes.c_int, ctypes.c_int, ctypes.c_void_p, ctypes.POINTER(ctypes.c_int),
                             ctypes.POINTER(ctypes.c_int)],
}, {"avl_clip_text_workspace_bytes": ctypes.c_longlong})

WIDTH, HEADS, LAYERS, CONTEXT, VOCAB, EMBED = 512, 8, 12, 77, 49408, 512


def clip_param_keys(layers=LAYERS):
    keys = ["token_embedding.weight", "positional_embedding"]
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        keys += [p + "ln_1.weight", p + "ln_1.bias", p + "attn.in_proj_weight", p + "attn.in_proj_bias",
                 p + "attn.out_proj.weight", p + "attn.out_proj.bias", p + "ln_2.weight", p + "ln_2.bias",
                 p + "mlp.c_fc.weight", p + "mlp.c_fc.bias", p + "mlp.c_proj.weight", p + "mlp.c_proj.bias"]
    return keys + ["ln_final.weight", "ln_final.bias", "text_projection"]


CLIP_PARAM_KEYS = clip_param_keys()


class _ResBlock(nn.Module):  # parameter container only (openai/CLIP ResidualAttentionBlock names)
    def __init__(self, d_model, n_head):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", nn.Identity()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.width, self.layers = width, layers
        self.resblocks = nn.Sequential(*[_ResBlock(width, heads) for _ in range(layers)])


class CLIPTextTower(nn.Module):
    def __init__(self, dedupe_zero_rows=True, chunk=512, layers=LAYERS):
        super().__init__()
        self.context_length, self.vocab_size, self.layers = CONTEXT, VOCAB, layers
        self._keys = clip_param_keys(layers)
        self.transformer = _Transformer(WIDTH, layers, HEADS)
        self.token_embedding = nn.Embedding(VOCAB, WIDTH)
        self.positional_embedding = nn.Parameter(torch.empty(CONTEXT, WIDTH))
        self.ln_final = nn.LayerNorm(WIDTH)
        self.text_projection = nn.Parameter(torch.empty(WIDTH, EMBED))
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592)
        self.dedupe_zero_rows, self.chunk = dedupe_zero_rows, chunk
        self._ptrs, self._ptr_key, self._ws = None, None, None
        # fp16 copies of the four linear weights per layer for the kind::f16 tensor-core path (the dtype `clip.load` gives
        # the tower on CUDA in the reference, policy.py:761); rebuilt when a weight's address or version changes
        self.half_gemms = True
        self._w16, self._w16_key, self._ptrs16 = None, None, None
        self.initialize_parameters()
        for p in self.parameters():
            p.requires_grad = False

    def initialize_parameters(self):  # openai/CLIP model.py initialize_parameters (text part)
        nn.init.normal_(self.token_embedding.weight, std=0.02)
        nn.init.normal_(self.positional_embedding, std=0.01)
        proj_std = (WIDTH ** -0.5) * ((2 * self.layers) ** -0.5)
        attn_std, fc_std = WIDTH ** -0.5, (2 * WIDTH) ** -0.5
        for block in self.transformer.resblocks:
            nn.init.normal_(block.attn.in_proj_weight, std=attn_std)
            nn.init.normal_(block.attn.out_proj.weight, std=proj_std)
            nn.init.normal_(block.mlp.c_fc.weight, std=fc_std)
            nn.init.normal_(block.mlp.c_proj.weight, std=proj_std)
        nn.init.normal_(self.text_projection, std=WIDTH ** -0.5)

    def _table(self):
        sd = dict(self.named_parameters())
        ts = [sd[k] for k in self._keys]
        key = tuple(t.data_ptr() for t in ts)
        if key != self._ptr_key:
            for t in ts:
                if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
                    raise _lib.AvlenError("CLIP text parameters must be contiguous fp32 CUDA tensors (call .float())")
            self._ptrs = (ctypes.c_void_p * len(key))(*key)
            self._ptr_key = key
        return self._ptrs

    _W16_SUFFIXES = ("attn.in_proj_weight", "attn.out_proj.weight", "mlp.c_fc.weight", "mlp.c_proj.weight")

    def _table16(self):
        sd = dict(self.named_parameters())
        ws = [sd[f"transformer.resblocks.{i}.{sfx}"] for i in range(self.layers) for sfx in self._W16_SUFFIXES]
        key = tuple((w.data_ptr(), w._version) for w in ws)
        if key != self._w16_key:
            self._w16 = [w.detach().half().contiguous() for w in ws]
            self._ptrs16 = (ctypes.c_void_p * len(ws))(*[w.data_ptr() for w in self._w16])
            self._w16_key = key
        return self._ptrs16

    @torch.no_grad()
    def encode_text(self, text):
        """text (B, L<=77) integer tokens -> (B, 512) fp32 (clip/model.py encode_text)."""
        text = text.to(torch.int64).contiguous()
        B, L = text.shape
        out = torch.empty((B, EMBED), device=text.device, dtype=torch.float32)
        tab = ctypes.cast(self._table(), ctypes.c_void_p)
        for b0 in range(0, B, self.chunk):
            n = min(self.chunk, B - b0)
            nbytes = int(_lib.lib().avl_clip_text_workspace_bytes(n, L))
            if self._ws is None or self._ws.numel() < nbytes:
                self._ws = torch.empty(nbytes, dtype=torch.uint8, device=text.device)
            if self.half_gemms:  # (the library still takes the fp32 path below 512 token rows or with tensor cores off)
                _lib.call("avl_clip_text_forward_f16", n, L, self.vocab_size, self.layers, text[b0:b0 + n].data_ptr(), tab,
                          ctypes.cast(self._table16(), ctypes.c_void_p), out[b0:b0 + n].data_ptr(), self._ws.data_ptr(),
                          int(self.dedupe_zero_rows), _lib.stream())
            else:
                _lib.call("avl_clip_text_forward", n, L, self.vocab_size, self.layers, text[b0:b0 + n].data_ptr(), tab,
                          out[b0:b0 + n].data_ptr(), self._ws.data_ptr(), int(self.dedupe_zero_rows), _lib.stream())
            self._last = (n, L)
        return out

    @torch.no_grad()
    def encode_text_cached(self, text):
        """``encode_text`` for the rollout, where row i is env i at every call: a dialog changes only when a query
        fires (ppo_trainer.py:548-553) and stays for NUM_DIALOG_STEPS steps, yet the reference runs the whole tower on
        all N rows every step (:625-637).  Rows whose 77 tokens equal the previous call's reuse that call's embedding;
        the others are encoded (unchanged rows are presented to the kernel as all-zero rows, which its device-side
        compaction folds into one shared sequence — no host synchronisation, no data-dependent launch shapes).
        Identical to ``encode_text`` as long as the (frozen) CLIP parameters do not change; ``reset_cache()`` after
        loading new ones."""
        text = text.to(torch.int64).contiguous()
        c = self.__dict__.get("_row_cache")
        key = (tuple(text.shape), text.device, self._ptr_key)
        if c is None or c[0] != key:
            emb = self.encode_text(text)
            # persistent buffers, updated in place: a rollout step captured into a CUDA graph keeps reading / writing them
            self.__dict__["_row_cache"] = (key, text.clone(), emb.clone())
            return emb
        _, old_text, old_emb = c
        changed = (text != old_text).any(dim=1, keepdim=True)
        emb_new = self.encode_text(torch.where(changed, text, torch.zeros_like(text)))
        emb = torch.where(changed, emb_new, old_emb)
        old_text.copy_(text)
        old_emb.copy_(emb)
        return emb

    def reset_cache(self):
        self.__dict__.pop("_row_cache", None)

    def last_counts(self):
        """(synchronising) distinct sequences / token rows processed by the last chunk."""
        a, b = ctypes.c_int(0), ctypes.c_int(0)
        n, L = self._last
        _lib.check(_lib.lib().avl_clip_text_status(n, L, self._ws.data_ptr(), ctypes.byref(a), ctypes.byref(b)))
        return a.value, b.value

    def forward(self, text):
        return self.encode_text(text)

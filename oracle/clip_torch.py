"""TEST INFRASTRUCTURE (oracle): CPU restatement of the CLIP ViT-B/32 text tower that AVLEN's dialog policy calls
(/root/reference/ss_baselines/savi/ppo/policy.py:761-762 ``clip.load("ViT-B/32")``, :847-849 ``encode_text``).

The algorithm lives in a third-party dependency that is ABSENT from the reference tree: openai/CLIP, installed from
git UNPINNED (/root/reference/README.md:61).  This file restates its published ``clip/model.py`` (``QuickGELU``,
``ResidualAttentionBlock``, ``Transformer``, ``CLIP.build_attention_mask`` / ``encode_text``) in plain PyTorch with
the same parameter names.  The reference holds no test / golden vector for it => **parity unpinned** at this
boundary; it is cross-checked against an independent implementation of the same architecture
(``transformers.CLIPTextModelWithProjection`` with the weights mapped, tests/test_oracle_clip.py).
Only tests/, smoke() and bench.py's CPU legs may import this module."""
from __future__ import annotations

from collections import OrderedDict

import torch
import torch.nn as nn


class QuickGELU(nn.Module):
    def forward(self, x):
        return x * torch.sigmoid(1.702 * x)


class ResidualAttentionBlock(nn.Module):
    def __init__(self, d_model, n_head):
        super().__init__()
        self.attn = nn.MultiheadAttention(d_model, n_head)
        self.ln_1 = nn.LayerNorm(d_model)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", nn.Linear(d_model, d_model * 4)), ("gelu", QuickGELU()),
                                              ("c_proj", nn.Linear(d_model * 4, d_model))]))
        self.ln_2 = nn.LayerNorm(d_model)

    def forward(self, x, attn_mask):
        y = self.ln_1(x)
        x = x + self.attn(y, y, y, need_weights=False, attn_mask=attn_mask)[0]
        return x + self.mlp(self.ln_2(x))


class _Transformer(nn.Module):
    def __init__(self, width, layers, heads):
        super().__init__()
        self.resblocks = nn.Sequential(*[ResidualAttentionBlock(width, heads) for _ in range(layers)])


class CLIPText(nn.Module):
    def __init__(self, width=512, heads=8, layers=12, context=77, vocab=49408, embed=512):
        super().__init__()
        self.transformer = _Transformer(width, layers, heads)
        self.token_embedding = nn.Embedding(vocab, width)
        self.positional_embedding = nn.Parameter(torch.randn(context, width) * 0.01)
        self.ln_final = nn.LayerNorm(width)
        self.text_projection = nn.Parameter(torch.randn(width, embed) * width ** -0.5)
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592)

    def encode_text(self, text):
        L = text.shape[1]
        x = self.token_embedding(text) + self.positional_embedding[:L]
        x = x.permute(1, 0, 2)  # NLD -> LND
        mask = torch.full((L, L), float("-inf")).triu_(1)  # build_attention_mask: causal
        for blk in self.transformer.resblocks:
            x = blk(x, mask)
        x = self.ln_final(x.permute(1, 0, 2))
        return x[torch.arange(x.shape[0]), text.argmax(dim=-1)] @ self.text_projection

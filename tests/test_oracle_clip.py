"""Pins the CLIP text-tower oracle (oracle/clip_torch.py, a restatement of openai/CLIP's encode_text) against an
independent implementation of the same architecture: transformers.CLIPTextModelWithProjection (quick_gelu, EOS pooled
at the largest token id) with the weights mapped across.  The reference itself holds no test for this dependency."""
import pytest
import torch

from oracle import clip_torch


def _hf_from_oracle(o, layers):
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    cfg = CLIPTextConfig(vocab_size=49408, hidden_size=512, intermediate_size=2048, num_hidden_layers=layers,
                         num_attention_heads=8, max_position_embeddings=77, hidden_act="quick_gelu", projection_dim=512,
                         eos_token_id=49407, bos_token_id=49406, pad_token_id=0)
    hf = CLIPTextModelWithProjection(cfg).eval()
    sd = {}
    osd = o.state_dict()
    sd["text_model.embeddings.token_embedding.weight"] = osd["token_embedding.weight"]
    sd["text_model.embeddings.position_embedding.weight"] = osd["positional_embedding"]
    for i in range(layers):
        p, q = f"transformer.resblocks.{i}.", f"text_model.encoder.layers.{i}."
        w, b = osd[p + "attn.in_proj_weight"], osd[p + "attn.in_proj_bias"]
        for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
            sd[q + f"self_attn.{n}.weight"], sd[q + f"self_attn.{n}.bias"] = w[j * 512:(j + 1) * 512], b[j * 512:(j + 1) * 512]
        sd[q + "self_attn.out_proj.weight"], sd[q + "self_attn.out_proj.bias"] = osd[p + "attn.out_proj.weight"], osd[p + "attn.out_proj.bias"]
        sd[q + "layer_norm1.weight"], sd[q + "layer_norm1.bias"] = osd[p + "ln_1.weight"], osd[p + "ln_1.bias"]
        sd[q + "layer_norm2.weight"], sd[q + "layer_norm2.bias"] = osd[p + "ln_2.weight"], osd[p + "ln_2.bias"]
        sd[q + "mlp.fc1.weight"], sd[q + "mlp.fc1.bias"] = osd[p + "mlp.c_fc.weight"], osd[p + "mlp.c_fc.bias"]
        sd[q + "mlp.fc2.weight"], sd[q + "mlp.fc2.bias"] = osd[p + "mlp.c_proj.weight"], osd[p + "mlp.c_proj.bias"]
    sd["text_model.final_layer_norm.weight"], sd["text_model.final_layer_norm.bias"] = osd["ln_final.weight"], osd["ln_final.bias"]
    sd["text_projection.weight"] = osd["text_projection"].t().contiguous()
    missing, unexpected = hf.load_state_dict(sd, strict=False)
    assert not unexpected and all("position_ids" in m for m in missing), (missing, unexpected)
    return hf


def test_clip_text_oracle_matches_transformers():
    pytest.importorskip("transformers")
    torch.manual_seed(0)
    layers = 3  # same block repeated; 3 layers keep the CPU test fast
    o = clip_torch.CLIPText(layers=layers).eval()
    for p in o.parameters():
        if p.dim() >= 2:
            torch.nn.init.normal_(p, std=0.05)
    hf = _hf_from_oracle(o, layers)
    tokens = torch.zeros(4, 77, dtype=torch.long)
    for b, n in enumerate((5, 11, 20, 1)):
        tokens[b, 0] = 49406
        tokens[b, 1:1 + n] = torch.randint(1, 49000, (n,))
        tokens[b, 1 + n] = 49407
    with torch.no_grad():
        ref = hf(input_ids=tokens).text_embeds
        out = o.encode_text(tokens)
    assert float((out - ref).abs().max()) < 2e-4 * max(1.0, float(ref.abs().max()))

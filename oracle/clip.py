"""TEST INFRASTRUCTURE — fp32 restatement of the OpenCLIP text encoder as used by FrozenOpenCLIPEmbedder
(terediff/model/clip.py:37-55; residual blocks terediff/model/open_clip/transformer.py:199-254; causal mask
model.py `build_attention_mask`).  Evaluated functionally from the embedder's state_dict (keys ``model.*``)."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def encode_tokens(sd: SD, tokens: torch.Tensor, heads: int = 16, layer: str = "penultimate") -> torch.Tensor:
    x = sd["model.token_embedding.weight"][tokens] + sd["model.positional_embedding"]
    B, L, E = x.shape
    n = 0
    while f"model.transformer.resblocks.{n}.ln_1.weight" in sd:
        n += 1
    mask = torch.full((L, L), float("-inf"), device=x.device).triu_(1)
    for i in range(n - (1 if layer == "penultimate" else 0)):
        p = f"model.transformer.resblocks.{i}"
        h = F.layer_norm(x, (E,), sd[p + ".ln_1.weight"], sd[p + ".ln_1.bias"], 1e-5)
        qkv = F.linear(h, sd[p + ".attn.in_proj_weight"], sd[p + ".attn.in_proj_bias"])
        q, k, v = (t.view(B, L, heads, E // heads).transpose(1, 2) for t in qkv.chunk(3, dim=-1))
        a = torch.softmax(q @ k.transpose(-1, -2) * (E // heads) ** -0.5 + mask, dim=-1) @ v
        x = x + F.linear(a.transpose(1, 2).reshape(B, L, E), sd[p + ".attn.out_proj.weight"], sd[p + ".attn.out_proj.bias"])
        h = F.layer_norm(x, (E,), sd[p + ".ln_2.weight"], sd[p + ".ln_2.bias"], 1e-5)
        h = F.linear(F.gelu(F.linear(h, sd[p + ".mlp.c_fc.weight"], sd[p + ".mlp.c_fc.bias"])),
                     sd[p + ".mlp.c_proj.weight"], sd[p + ".mlp.c_proj.bias"])
        x = x + h
    return F.layer_norm(x, (E,), sd["model.ln_final.weight"], sd["model.ln_final.bias"], 1e-5)

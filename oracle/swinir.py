"""TEST INFRASTRUCTURE — fp32 restatement of the SwinIR cleaner as configured for TeReDiff
(terediff/model/swinir.py:624-892 with configs/val/val_terediff.yaml:69-85: unshuffle 8, embed 180, 8 RSTB x 6 blocks,
6 heads, window 8, mlp_ratio 2, 'nearest+conv' x8, '1conv').  Evaluated functionally from the module's state_dict."""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
WS = 8
RGB_MEAN = (0.4488, 0.4371, 0.4040)


def _windows(x, ws):           # swinir.py:37-50  (B,H,W,C) -> (B*nW, ws*ws, C)
    B, H, W, C = x.shape
    return x.view(B, H // ws, ws, W // ws, ws, C).permute(0, 1, 3, 2, 4, 5).reshape(-1, ws * ws, C)


def _unwindows(w, ws, H, W):   # swinir.py:53-66
    B = w.shape[0] // (H * W // ws // ws)
    return w.view(B, H // ws, W // ws, ws, ws, -1).permute(0, 1, 3, 2, 4, 5).reshape(B, H, W, -1)


def shift_mask(H, W, ws=WS, shift=WS // 2):   # swinir.py:222-243
    img = torch.zeros((1, H, W, 1))
    cnt = 0
    for hs in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
        for wsl in (slice(0, -ws), slice(-ws, -shift), slice(-shift, None)):
            img[:, hs, wsl, :] = cnt
            cnt += 1
    m = _windows(img, ws).view(-1, ws * ws)
    d = m.unsqueeze(1) - m.unsqueeze(2)
    return torch.where(d != 0, torch.full_like(d, -100.0), torch.zeros_like(d))   # (nW, N, N)


def rel_index(ws: int = WS) -> torch.Tensor:   # swinir.py:97-109: a constant of the architecture (N, N) long
    coords = torch.stack(torch.meshgrid(torch.arange(ws), torch.arange(ws), indexing="ij")).flatten(1)
    rel = (coords[:, :, None] - coords[:, None, :]).permute(1, 2, 0).contiguous()
    rel[:, :, 0] += ws - 1
    rel[:, :, 1] += ws - 1
    rel[:, :, 0] *= 2 * ws - 1
    return rel.sum(-1)


def rel_bias(sd: SD, p: str, heads: int) -> torch.Tensor:   # swinir.py:129-132 -> (heads, N, N)
    idx = rel_index().view(-1).to(sd[p + ".relative_position_bias_table"].device)
    N = WS * WS
    return sd[p + ".relative_position_bias_table"][idx].view(N, N, heads).permute(2, 0, 1)


def block(sd: SD, p: str, x, H, W, heads, shift):   # swinir.py:245-288
    B, L, C = x.shape
    h = F.layer_norm(x, (C,), sd[p + ".norm1.weight"], sd[p + ".norm1.bias"], 1e-5).view(B, H, W, C)
    if shift:
        h = torch.roll(h, shifts=(-shift, -shift), dims=(1, 2))
    w = _windows(h, WS)
    qkv = F.linear(w, sd[p + ".attn.qkv.weight"], sd[p + ".attn.qkv.bias"])
    Bw, N, _ = w.shape
    q, k, v = qkv.view(Bw, N, 3, heads, C // heads).permute(2, 0, 3, 1, 4)
    attn = (q * (C // heads) ** -0.5) @ k.transpose(-2, -1) + rel_bias(sd, p + ".attn", heads).unsqueeze(0)
    if shift:
        m = shift_mask(H, W).to(x.device)
        attn = (attn.view(Bw // m.shape[0], m.shape[0], heads, N, N) + m[None, :, None]).view(-1, heads, N, N)
    o = (torch.softmax(attn, -1) @ v).transpose(1, 2).reshape(Bw, N, C)
    o = F.linear(o, sd[p + ".attn.proj.weight"], sd[p + ".attn.proj.bias"])
    o = _unwindows(o, WS, H, W)
    if shift:
        o = torch.roll(o, shifts=(shift, shift), dims=(1, 2))
    x = x + o.view(B, L, C)
    h = F.layer_norm(x, (C,), sd[p + ".norm2.weight"], sd[p + ".norm2.bias"], 1e-5)
    h = F.linear(F.gelu(F.linear(h, sd[p + ".mlp.fc1.weight"], sd[p + ".mlp.fc1.bias"])), sd[p + ".mlp.fc2.weight"], sd[p + ".mlp.fc2.bias"])
    return x + h


def swinir_forward(sd: SD, x: torch.Tensor, heads: int = 6, sf: int = 8) -> torch.Tensor:
    """x (B,3,H,W) in [0,1], H and W multiples of 64 -> cleaned image (B,3,H,W)  (swinir.py:856-892)."""
    conv = lambda t, n: F.conv2d(t, sd[n + ".weight"], sd[n + ".bias"], padding=1)
    mean = torch.tensor(RGB_MEAN, device=x.device).view(1, 3, 1, 1)
    x = x - mean
    f0 = conv(F.pixel_unshuffle(x, sf), "conv_first.1")
    B, C, H, W = f0.shape
    t = f0.flatten(2).transpose(1, 2)
    t = F.layer_norm(t, (C,), sd["patch_embed.norm.weight"], sd["patch_embed.norm.bias"], 1e-5)
    nl = 0
    while f"layers.{nl}.conv.weight" in sd:
        nl += 1
    for i in range(nl):
        r = t
        j = 0
        while f"layers.{i}.residual_group.blocks.{j}.norm1.weight" in sd:
            r = block(sd, f"layers.{i}.residual_group.blocks.{j}", r, H, W, heads, 0 if j % 2 == 0 else WS // 2)
            j += 1
        r = conv(r.transpose(1, 2).view(B, C, H, W), f"layers.{i}.conv").flatten(2).transpose(1, 2)
        t = t + r
    t = F.layer_norm(t, (C,), sd["norm.weight"], sd["norm.bias"], 1e-5)
    y = conv(t.transpose(1, 2).view(B, C, H, W), "conv_after_body") + f0
    y = F.leaky_relu(conv(y, "conv_before_upsample.0"), 0.01)
    for n in ("conv_up1", "conv_up2", "conv_up3"):
        y = F.leaky_relu(conv(F.interpolate(y, scale_factor=2, mode="nearest"), n), 0.2)
    y = conv(F.leaky_relu(conv(y, "conv_hr"), 0.2), "conv_last")
    return y + mean

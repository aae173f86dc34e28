"""TEST INFRASTRUCTURE — fp32 restatement of the AutoencoderKL *decoder* (terediff/model/vae.py:429-582) and of the
PSNR formula (terediff/utils/common.py:361-392).  Used only to turn final latents into images for the north-star
"final image PSNR >= 40 dB" gate; the VAE itself is a SURVEY §8f "next" component of the product.

Evaluated functionally from ``AutoencoderKL.state_dict()`` keys (``post_quant_conv.*``, ``decoder.*``); block structure
is recovered from the keys present.  Pinned against the imported reference by tests/golden/vae_decode.npz.
"""
from __future__ import annotations

from typing import Dict

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]


def _conv(sd, p, x, pad=None):
    w = sd[p + ".weight"]
    return F.conv2d(x, w, sd[p + ".bias"], padding=w.shape[-1] // 2 if pad is None else pad)


def _gn(sd, p, x):
    return F.group_norm(x, 32, sd[p + ".weight"], sd[p + ".bias"], 1e-6)        # Normalize, vae.py:18-21


def _swish(x):
    return x * torch.sigmoid(x)                                                 # nonlinearity, vae.py:13-15


def _resblock(sd, p, x):
    """ResnetBlock.forward with temb=None — vae.py:97-121."""
    h = _conv(sd, p + ".conv1", _swish(_gn(sd, p + ".norm1", x)))
    h = _conv(sd, p + ".conv2", _swish(_gn(sd, p + ".norm2", h)))
    if (p + ".nin_shortcut.weight") in sd:
        x = _conv(sd, p + ".nin_shortcut", x)
    elif (p + ".conv_shortcut.weight") in sd:
        x = _conv(sd, p + ".conv_shortcut", x)
    return x + h


def _attn(sd, p, x):
    """Single-head spatial self-attention over C channels — SDPAttnBlock.forward, vae.py:253-281."""
    h = _gn(sd, p + ".norm", x)
    B, C, H, W = h.shape
    q, k, v = (_conv(sd, f"{p}.{n}", h).flatten(2).transpose(1, 2) for n in ("q", "k", "v"))
    a = torch.softmax(q @ k.transpose(1, 2) * C ** -0.5, dim=-1) @ v
    return x + _conv(sd, p + ".proj_out", a.transpose(1, 2).reshape(B, C, H, W))


def vae_decode(sd: SD, z: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.decode — vae.py:579-582 -> Decoder.forward :526-559."""
    D = "decoder"
    h = _conv(sd, D + ".conv_in", _conv(sd, "post_quant_conv", z))
    h = _resblock(sd, D + ".mid.block_1", h)
    h = _attn(sd, D + ".mid.attn_1", h)
    h = _resblock(sd, D + ".mid.block_2", h)
    levels = 0
    while any(k.startswith(f"{D}.up.{levels}.") for k in sd):
        levels += 1
    for lvl in reversed(range(levels)):
        i = 0
        while (f"{D}.up.{lvl}.block.{i}.norm1.weight") in sd:
            h = _resblock(sd, f"{D}.up.{lvl}.block.{i}", h)
            if (f"{D}.up.{lvl}.attn.{i}.norm.weight") in sd:
                h = _attn(sd, f"{D}.up.{lvl}.attn.{i}", h)
            i += 1
        if (f"{D}.up.{lvl}.upsample.conv.weight") in sd:
            h = _conv(sd, f"{D}.up.{lvl}.upsample.conv", F.interpolate(h, scale_factor=2.0, mode="nearest"))
    return _conv(sd, D + ".conv_out", _swish(_gn(sd, D + ".norm_out", h)))


def vae_encode_moments(sd: SD, x: torch.Tensor) -> torch.Tensor:
    """AutoencoderKL.encode up to the posterior parameters — vae.py:573-577, Encoder.forward :399-427,
    Downsample.forward :50-57 (zero pad right/bottom, stride-2 conv without padding)."""
    E = "encoder"
    h = _conv(sd, E + ".conv_in", x)
    lvl = 0
    while any(k.startswith(f"{E}.down.{lvl}.") for k in sd):
        i = 0
        while (f"{E}.down.{lvl}.block.{i}.norm1.weight") in sd:
            h = _resblock(sd, f"{E}.down.{lvl}.block.{i}", h)
            i += 1
        if (f"{E}.down.{lvl}.downsample.conv.weight") in sd:
            h = F.conv2d(F.pad(h, (0, 1, 0, 1)), sd[f"{E}.down.{lvl}.downsample.conv.weight"],
                         sd[f"{E}.down.{lvl}.downsample.conv.bias"], stride=2)
        lvl += 1
    h = _resblock(sd, E + ".mid.block_1", h)
    h = _attn(sd, E + ".mid.attn_1", h)
    h = _resblock(sd, E + ".mid.block_2", h)
    h = _conv(sd, E + ".conv_out", _swish(_gn(sd, E + ".norm_out", h)))
    return _conv(sd, "quant_conv", h)


def image_to_latent(sd: SD, img01: torch.Tensor, scale_factor: float = 0.18215) -> torch.Tensor:
    """prepare_condition's c_img (cldm.py:143-158): mode of the posterior of (img*2-1), times the latent scale."""
    return vae_encode_moments(sd, img01 * 2 - 1)[:, :4] * scale_factor


def latent_to_image(sd: SD, z: torch.Tensor, scale_factor: float = 0.18215) -> torch.Tensor:
    """ControlLDM.vae_decode + the clamp of val_patches.py:369: clamp((decode(z / s) + 1) / 2, 0, 1)."""
    return ((vae_decode(sd, z / scale_factor) + 1) / 2).clamp(0, 1)


def psnr(img: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
    """calculate_psnr_pt with crop_border=0 — common.py:388-392 (images in [0,1])."""
    mse = torch.mean((img.double() - img2.double()) ** 2, dim=[1, 2, 3])
    return 10.0 * torch.log10(1.0 / (mse + 1e-8))

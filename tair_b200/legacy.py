"""The legacy DiffBIR-style surface around the denoiser: ``make_tiled_fn`` latent tiling and ``Pipeline.apply_cldm/run``
(terediff/utils/common.py:125-234, terediff/pipeline.py:25-397) — SURVEY.md §8f rank 4.

In the reference this surface no longer works with ``SpacedSampler``: ``sample`` was changed to return
``(x, sampled_unet_feats)`` (spaced_sampler.py:243) while ``Pipeline.apply_cldm`` still slices its result as a tensor
(pipeline.py:218-233), and the spaced sampler ignores ``tiled`` (spaced_sampler.py:192-243 never calls
``make_tiled_fn``).  Here both are repaired: ``apply_cldm`` unpacks the tuple, and ``SpacedSampler.sample(tiled=True)``
wraps the model in ``make_tiled_fn`` the way the reference's other samplers do (ddim_sampler.py:165-180).

Everything in this file is host-side orchestration; the networks it drives run on the sm_100a kernels.  The tile
accumulation (``out += fn(tile) * w; out / count``) and the colour-fix post-processing are O(image) torch elementwise /
depth-wise-blur plumbing on fp32 tensors, once per image, outside the denoising loop.

Geometry note: the implicit-GEMM convolution tiles output rows in 128-pixel M tiles, so latent widths must divide 128 or
be a multiple of 128 at every UNet level.  512-pixel diffusion tiles (64-wide latents) always qualify, so ``apply_cldm``
switches latent tiling on by itself when the padded latent does not.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple

import numpy as np
import torch
from torch.nn import functional as F


# ---- terediff/utils/common.py:125-234 ------------------------------------------------------------------------------

def sliding_windows(h: int, w: int, tile_size: int, tile_stride: int) -> List[Tuple[int, int, int, int]]:
    """common.py:125-141: top-left aligned windows plus one flush with the far edge when the stride does not fit."""
    his = list(range(0, h - tile_size + 1, tile_stride))
    if (h - tile_size) % tile_stride != 0:
        his.append(h - tile_size)
    wis = list(range(0, w - tile_size + 1, tile_stride))
    if (w - tile_size) % tile_stride != 0:
        wis.append(w - tile_size)
    return [(hi, hi + tile_size, wi, wi + tile_size) for hi in his for wi in wis]


def gaussian_weights(tile_width: int, tile_height: int) -> np.ndarray:
    """common.py:145-172 (note the asymmetric midpoints: (W-1)/2 for x, H/2 for y — kept as in the reference)."""
    var = 0.01
    mx = (tile_width - 1) / 2
    xs = np.arange(tile_width, dtype=np.float64)
    x_probs = np.exp(-(xs - mx) * (xs - mx) / (tile_width * tile_width) / (2 * var)) / np.sqrt(2 * np.pi * var)
    my = tile_height / 2
    ys = np.arange(tile_height, dtype=np.float64)
    y_probs = np.exp(-(ys - my) * (ys - my) / (tile_height * tile_height) / (2 * var)) / np.sqrt(2 * np.pi * var)
    return np.outer(y_probs, x_probs)


def make_tiled_fn(fn: Callable, size: int, stride: int, scale_type: str = "up", scale: int = 1,
                  channel: Optional[int] = None, weight: str = "gaussian", dtype: Optional[torch.dtype] = None,
                  device: Optional[torch.device] = None, progress: bool = True) -> Callable:
    """common.py:175-234.  Splits only the first input; when extra arguments are given the tile bounds are passed to
    ``fn`` as ``hi, hi_end, wi, wi_end`` keyword arguments (so the caller can slice its conditioning)."""
    def tiled_fn(x: torch.Tensor, *args, **kwargs) -> torch.Tensor:
        scale_fn = (lambda n: int(n * scale)) if scale_type == "up" else (lambda n: int(n // scale))
        b, c, h, w = x.size()
        out_dtype, out_device = dtype or x.dtype, device or x.device
        out = torch.zeros((b, channel or c, scale_fn(h), scale_fn(w)), dtype=out_dtype, device=out_device)
        count = torch.zeros_like(out, dtype=torch.float32)
        ws = scale_fn(size)
        wts = gaussian_weights(ws, ws)[None, None] if weight == "gaussian" else np.ones((1, 1, ws, ws))
        wts = torch.tensor(wts, dtype=out_dtype, device=out_device)
        for hi, hi_end, wi, wi_end in sliding_windows(h, w, size, stride):
            x_tile = x[..., hi:hi_end, wi:wi_end]
            ohi, ohi_end, owi, owi_end = map(scale_fn, (hi, hi_end, wi, wi_end))
            if len(args) or len(kwargs):
                kwargs.update(dict(hi=hi, hi_end=hi_end, wi=wi, wi_end=wi_end))
            out[..., ohi:ohi_end, owi:owi_end] += fn(x_tile, *args, **kwargs) * wts
            count[..., ohi:ohi_end, owi:owi_end] += wts
        return out / count
    return tiled_fn


def tiled_model(model: Callable, tile_size: int, tile_stride: int) -> Callable:
    """``model(x, t, cond) -> (eps, feats)`` evaluated tile by tile on the latent with ``c_img`` sliced alongside
    (ddim_sampler.py:165-180); decoder features do not tile, so the wrapped model returns ``(eps, None)``."""
    inner = make_tiled_fn(
        lambda x_tile, t, cond, hi, hi_end, wi, wi_end: model(
            x_tile.contiguous(), t, {"c_txt": cond["c_txt"], "c_img": cond["c_img"][..., hi:hi_end, wi:wi_end].contiguous()})[0],
        tile_size, tile_stride, progress=False)
    return lambda x, t, cond: (inner(x, t, cond), None)


# ---- terediff/utils/common.py:33-79 (colour fix) --------------------------------------------------------------------

def wavelet_blur(image: torch.Tensor, radius: int) -> torch.Tensor:
    k = torch.tensor([[0.0625, 0.125, 0.0625], [0.125, 0.25, 0.125], [0.0625, 0.125, 0.0625]], dtype=image.dtype,
                     device=image.device)[None, None].repeat(3, 1, 1, 1)
    image = F.pad(image, (radius, radius, radius, radius), mode="replicate")
    return F.conv2d(image, k, groups=3, dilation=radius)


def wavelet_decomposition(image: torch.Tensor, levels: int = 5):
    high = torch.zeros_like(image)
    low = image
    for i in range(levels):
        low = wavelet_blur(image, 2 ** i)
        high += image - low
        image = low
    return high, low


def wavelet_reconstruction(content: torch.Tensor, style: torch.Tensor) -> torch.Tensor:
    """common.py:68-79: high frequencies of the sample, low frequencies (colour) of the cleaned condition image."""
    return wavelet_decomposition(content)[0] + wavelet_decomposition(style)[1]


# ---- terediff/pipeline.py --------------------------------------------------------------------------------------------

def resize_short_edge_to(imgs: torch.Tensor, size: int) -> torch.Tensor:
    """pipeline.py:25-34."""
    _, _, h, w = imgs.size()
    if h == w:
        oh, ow = size, size
    elif h < w:
        oh, ow = size, int(w * (size / h))
    else:
        oh, ow = int(h * (size / w)), size
    return F.interpolate(imgs, size=(oh, ow), mode="bicubic", antialias=True)


def pad_to_multiples_of(imgs: torch.Tensor, multiple: int) -> torch.Tensor:
    """pipeline.py:37-42."""
    _, _, h, w = imgs.size()
    if h % multiple == 0 and w % multiple == 0:
        return imgs.clone()
    ph, pw = ((v + multiple - 1) // multiple * multiple - v for v in (h, w))
    return F.pad(imgs, pad=(0, pw, 0, ph), mode="constant", value=0)


def _latent_ok(n: int) -> bool:
    """Every UNet level (n, n/2, n/4, n/8) must divide the 128-row M tile or be a multiple of it."""
    return all((128 % (n >> k) == 0 or (n >> k) % 128 == 0) and (n >> k) > 0 and n % 8 == 0 for k in range(4))


class Pipeline:
    """pipeline.py:45-327 with the spaced sampler (the only sampler of the TeReDiff path; the DDIM / DPM-Solver / EDM
    samplers of pipeline.py:188-207 are out of scope)."""

    def __init__(self, cleaner, cldm, diffusion, cond_fn, device) -> None:
        self.cleaner, self.cldm, self.diffusion, self.cond_fn, self.device = cleaner, cldm, diffusion, cond_fn, device
        self.output_size: Optional[Tuple[int, int]] = None
        if cond_fn is not None:
            raise NotImplementedError("restoration guidance (utils/cond_fn.py) needs autograd through the VAE decoder: out of scope")

    def set_output_size(self, lq_size) -> None:
        self.output_size = tuple(lq_size[2:])

    def apply_cleaner(self, lq: torch.Tensor, tiled: bool, tile_size: int, tile_stride: int) -> torch.Tensor:
        raise NotImplementedError

    @torch.no_grad()
    def apply_cldm(self, cond_img, steps, strength, vae_encoder_tiled, vae_encoder_tile_size, vae_decoder_tiled,
                   vae_decoder_tile_size, cldm_tiled, cldm_tile_size, cldm_tile_stride, pos_prompt, neg_prompt, cfg_scale,
                   start_point_type, sampler_type, noise_aug, rescale_cfg, s_churn=0, s_tmin=0, s_tmax=0, s_noise=1,
                   eta=0, order=1) -> torch.Tensor:
        from .sampler import SpacedSampler
        if sampler_type != "spaced":
            raise NotImplementedError(f"sampler_type {sampler_type!r}: tair_b200 implements the spaced sampler of the TeReDiff path")
        if vae_encoder_tiled or vae_decoder_tiled:
            raise NotImplementedError("tiled VAE (utils/tilevae.py) is out of scope; the kernel VAE handles 512-pixel multiples untiled")
        bs, _, h0, w0 = cond_img.shape
        cond_img = pad_to_multiples_of(cond_img, multiple=64 if not cldm_tiled else 8)        # pipeline.py:100-104
        cond = self.cldm.prepare_condition(cond_img, [pos_prompt] * bs)
        uncond = self.cldm.prepare_condition(cond_img, [neg_prompt] * bs)
        h1, w1 = cond["c_img"].shape[2:]
        if cldm_tiled and (h1 < cldm_tile_size // 8 or w1 < cldm_tile_size // 8):            # pipeline.py:131-133
            cldm_tiled = False
        if not cldm_tiled:
            cond["c_img"] = pad_to_multiples_of(cond["c_img"], multiple=8)
            uncond["c_img"] = pad_to_multiples_of(uncond["c_img"], multiple=8)
            h2, w2 = cond["c_img"].shape[2:]
            if not (_latent_ok(h2) and _latent_ok(w2)):
                # geometry the convolution kernels do not tile: fall back to 512-pixel diffusion tiles (see module doc)
                cldm_tiled, cldm_tile_size, cldm_tile_stride = True, 512, 256
                if h2 < 64 or w2 < 64:
                    raise ValueError(f"latent {h2}x{w2} is smaller than one 64x64 diffusion tile")
        elif cldm_tile_size % 64 != 0:
            raise ValueError("Diffusion tile size must be a multiple of 64")
        h2, w2 = cond["c_img"].shape[2:]
        if start_point_type == "cond":                                                       # pipeline.py:148-160
            t_last = torch.full((bs,), self.diffusion.num_timesteps - 1, dtype=torch.long, device=self.device)
            x_T = self.diffusion.q_sample(cond["c_img"], t_last, torch.randn(cond["c_img"].shape, device=self.device))
        else:
            x_T = torch.randn((bs, 4, h2, w2), dtype=torch.float32, device=self.device)
        if noise_aug > 0:                                                                    # pipeline.py:163-169
            t_aug = torch.full((bs,), noise_aug, dtype=torch.long, device=self.device)
            cond["c_img"] = self.diffusion.q_sample(cond["c_img"], t_aug, torch.randn_like(cond["c_img"]))
            uncond["c_img"] = cond["c_img"].detach().clone()
        saved = self.cldm.control_scales
        self.cldm.control_scales = [strength] * 13                                           # pipeline.py:175-176
        try:
            sampler = SpacedSampler(self.diffusion.betas, self.diffusion.parameterization, rescale_cfg)
            # repaired against the sampler's (x, feats) return type (spaced_sampler.py:243)
            z, _ = sampler.sample(model=self.cldm, device=self.device, steps=steps, x_size=(bs, 4, h2, w2), cond=cond,
                                  uncond=uncond, cfg_scale=cfg_scale, tiled=cldm_tiled, tile_size=cldm_tile_size // 8,
                                  tile_stride=cldm_tile_stride // 8, x_T=x_T, progress=False,
                                  use_cuda_graph=not cldm_tiled)
        finally:
            self.cldm.control_scales = saved
        z = z[..., :h1, :w1]
        return self.cldm.vae_decode(z)[:, :, :h0, :w0]

    @torch.no_grad()
    def run(self, lq: np.ndarray, steps, strength, cleaner_tiled, cleaner_tile_size, cleaner_tile_stride,
            vae_encoder_tiled, vae_encoder_tile_size, vae_decoder_tiled, vae_decoder_tile_size, cldm_tiled, cldm_tile_size,
            cldm_tile_stride, pos_prompt, neg_prompt, cfg_scale, start_point_type, sampler_type, noise_aug, rescale_cfg,
            s_churn=0, s_tmin=0, s_tmax=0, s_noise=1, eta=0, order=1) -> np.ndarray:
        """pipeline.py:236-327: (n,H,W,3) uint8 -> (n,H',W',3) uint8."""
        lq_t = torch.tensor(lq, dtype=torch.float32, device=self.device).div(255).clamp(0, 1).permute(0, 3, 1, 2).contiguous()
        self.set_output_size(lq_t.size())
        cond_img = self.apply_cleaner(lq_t, cleaner_tiled, cleaner_tile_size, cleaner_tile_stride)
        assert all(v >= 512 for v in cond_img.shape[2:]), "stage-1 output must be at least 512 pixels on each side"
        sample = self.apply_cldm(cond_img, steps, strength, vae_encoder_tiled, vae_encoder_tile_size, vae_decoder_tiled,
                                 vae_decoder_tile_size, cldm_tiled, cldm_tile_size, cldm_tile_stride, pos_prompt,
                                 neg_prompt, cfg_scale, start_point_type, sampler_type, noise_aug, rescale_cfg, s_churn,
                                 s_tmin, s_tmax, s_noise, eta, order)
        sample = F.interpolate(wavelet_reconstruction((sample + 1) / 2, cond_img), size=self.output_size, mode="bicubic",
                               antialias=True)
        return (sample * 255.0).clamp(0, 255).to(torch.uint8).permute(0, 2, 3, 1).contiguous().cpu().numpy()


class SwinIRPipeline(Pipeline):
    """pipeline.py:369-397."""

    def apply_cleaner(self, lq: torch.Tensor, tiled: bool, tile_size: int, tile_stride: int) -> torch.Tensor:
        if tiled and (lq.size(2) < tile_size or lq.size(3) < tile_size):
            tiled = False
        if tiled and tile_size % 64 != 0:
            raise ValueError("SwinIR (cleaner) tile size must be a multiple of 64")
        if not tiled:
            if min(lq.shape[2:]) < 512:
                lq = resize_short_edge_to(lq, size=512)
            h0, w0 = lq.shape[2:]
            lq = pad_to_multiples_of(lq, multiple=64)
            return self.cleaner(lq)[:, :, :h0, :w0]
        out = make_tiled_fn(self.cleaner, size=tile_size, stride=tile_stride, progress=False)(lq)
        if min(out.shape[2:]) < 512:
            out = resize_short_edge_to(out, size=512)
        return out

"""AutoencoderKL (encoder + decoder) on the sm_100a kernels (SURVEY.md §8f "next", rank 1).

Mirrors terediff/model/vae.py — ``ResnetBlock`` (:60-121), ``SDPAttnBlock`` (:232-281), ``Upsample`` (:24-37),
``Decoder`` (:429-559), ``AutoencoderKL`` (:562-582) — with the same submodule names, so the reference checkpoint's
``decoder.*`` / ``post_quant_conv.*`` keys load unchanged.  The decoder runs once per tile after the 50 denoising steps
(2.5 TFLOP per tile): GroupNorm(eps 1e-6)+swish -> implicit-GEMM conv3x3 with the residual fused into the epilogue,
nearest x2 upsampling, and the single-head 512-wide mid attention evaluated as GEMM -> row softmax -> GEMM
(head_dim 512 does not fit the 64-wide flash kernel).  The encoder (``Encoder`` :284-427, ``Downsample`` :40-57 with its
asymmetric (0,1,0,1) zero padding done by TMA out-of-bounds fill) runs once per tile in ``prepare_condition``.
"""
from __future__ import annotations

from typing import List

import torch
from torch import nn

from .. import ops
from .util import BF16, Conv1x1, Conv3x3, GroupNorm


def Normalize(in_channels: int, num_groups: int = 32) -> GroupNorm:
    """vae.py:18-21."""
    return GroupNorm(num_groups, in_channels, eps=1e-6)


class ResnetBlock(nn.Module):
    """vae.py:60-121 with temb_channels == 0 (the decoder's setting)."""

    def __init__(self, in_channels: int, out_channels: int = None):
        super().__init__()
        out_channels = in_channels if out_channels is None else out_channels
        self.in_channels, self.out_channels = in_channels, out_channels
        self.norm1 = Normalize(in_channels)
        self.conv1 = Conv3x3(in_channels, out_channels)
        self.norm2 = Normalize(out_channels)
        self.dropout = nn.Dropout(0.0)
        self.conv2 = Conv3x3(out_channels, out_channels)
        if in_channels != out_channels:
            self.nin_shortcut = Conv1x1(in_channels, out_channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        h = self.conv1(self.norm1(x, act=ops.ACT_SILU))
        h = self.norm2(h, act=ops.ACT_SILU)
        skip = self.nin_shortcut(x) if self.in_channels != self.out_channels else x
        return self.conv2(h, residual=skip)


class AttnBlock(nn.Module):
    """SDPAttnBlock (vae.py:232-281): GroupNorm -> 1x1 q,k,v -> softmax(q k^T / sqrt(C)) v -> 1x1 proj + x."""

    def __init__(self, in_channels: int):
        super().__init__()
        self.in_channels = in_channels
        self.norm = Normalize(in_channels)
        self.q = Conv1x1(in_channels, in_channels)
        self.k = Conv1x1(in_channels, in_channels)
        self.v = Conv1x1(in_channels, in_channels)
        self.proj_out = Conv1x1(in_channels, in_channels)

    def _qkv(self):
        st = (self.q._stamp(), self.k._stamp(), self.v._stamp())
        if getattr(self, "_qkv_stamp", None) != st:
            with torch.no_grad():
                C = self.in_channels
                w = torch.cat([m.weight.detach().reshape(C, C) for m in (self.q, self.k, self.v)], 0)
                b = torch.cat([m.bias.detach() for m in (self.q, self.k, self.v)], 0)
                self._qkv_pack = (w.to(BF16).contiguous(), b.float().contiguous())
            self._qkv_stamp = st
        return self._qkv_pack

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        B, H, W, C = x.shape
        L = H * W
        w, b = self._qkv()
        qkv = ops.gemm(self.norm(x).view(B * L, C), w, bias=b).view(B, L, 3 * C)
        vt = ops.transpose(qkv[:, :, 2 * C:].contiguous())                 # [B, C, L]: V with the token index contiguous
        att = torch.empty((B * L, C), device=x.device, dtype=BF16)
        for i in range(B):                                                 # 32 MB score matrix per image, reused
            s = ops.gemm(qkv[i, :, :C], qkv[i, :, C:2 * C])                 # [L, L] = q k^T
            p = ops.softmax_rows(s, scale=C ** -0.5, out=s)
            ops.gemm(p, vt[i], out=att[i * L:(i + 1) * L])                  # [L, C] = p v
        return self.proj_out(att.view(B, H, W, C), residual=x)


class Upsample(nn.Module):
    """vae.py:24-37."""

    def __init__(self, in_channels: int, with_conv: bool = True):
        super().__init__()
        assert with_conv
        self.with_conv = with_conv
        self.conv = Conv3x3(in_channels, in_channels)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.conv(ops.upsample2x(x))


class Decoder(nn.Module):
    """vae.py:429-559."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, give_pre_end=False, tanh_out=False,
                 **ignorekwargs):
        super().__init__()
        if give_pre_end or tanh_out or len(attn_resolutions):
            raise NotImplementedError("tair_b200 VAE decoder covers the TeReDiff config (configs/val/val_terediff.yaml:21-37)")
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        block_in = ch * ch_mult[-1]
        self.conv_in = Conv3x3(z_channels, block_in)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.up = nn.ModuleList()
        for i_level in reversed(range(self.num_resolutions)):
            block = nn.ModuleList()
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks + 1):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            up = nn.Module()
            up.block = block
            up.attn = nn.ModuleList()
            if i_level != 0:
                up.upsample = Upsample(block_in, resamp_with_conv)
            self.up.insert(0, up)
        self.norm_out = Normalize(block_in)
        self.conv_out = Conv3x3(block_in, out_ch)

    def forward(self, z_nhwc: torch.Tensor) -> torch.Tensor:
        h = self.conv_in(z_nhwc)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        for i_level in reversed(range(self.num_resolutions)):
            for blk in self.up[i_level].block:
                h = blk(h)
            if i_level != 0:
                h = self.up[i_level].upsample(h)
        return self.conv_out(self.norm_out(h, act=ops.ACT_SILU))


class Downsample(nn.Module):
    """vae.py:40-57: F.pad(x, (0,1,0,1)) then conv3x3 stride 2 padding 0."""

    def __init__(self, in_channels: int, with_conv: bool = True):
        super().__init__()
        assert with_conv
        self.with_conv = with_conv
        self.conv = Conv3x3(in_channels, in_channels, stride=2, asymmetric_pad=True)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.conv(x)


class Encoder(nn.Module):
    """vae.py:284-427."""

    def __init__(self, *, ch, out_ch, ch_mult=(1, 2, 4, 8), num_res_blocks, attn_resolutions, dropout=0.0,
                 resamp_with_conv=True, in_channels, resolution, z_channels, double_z=True, **ignore_kwargs):
        super().__init__()
        if len(attn_resolutions):
            raise NotImplementedError("tair_b200 VAE encoder covers the TeReDiff config (attn_resolutions: [])")
        self.num_resolutions = len(ch_mult)
        self.num_res_blocks = num_res_blocks
        self.conv_in = Conv3x3(in_channels, ch)
        in_ch_mult = (1,) + tuple(ch_mult)
        self.down = nn.ModuleList()
        block_in = ch
        for i_level in range(self.num_resolutions):
            block = nn.ModuleList()
            block_in = ch * in_ch_mult[i_level]
            block_out = ch * ch_mult[i_level]
            for _ in range(num_res_blocks):
                block.append(ResnetBlock(block_in, block_out))
                block_in = block_out
            down = nn.Module()
            down.block = block
            down.attn = nn.ModuleList()
            if i_level != self.num_resolutions - 1:
                down.downsample = Downsample(block_in, resamp_with_conv)
            self.down.append(down)
        self.mid = nn.Module()
        self.mid.block_1 = ResnetBlock(block_in, block_in)
        self.mid.attn_1 = AttnBlock(block_in)
        self.mid.block_2 = ResnetBlock(block_in, block_in)
        self.norm_out = Normalize(block_in)
        self.conv_out = Conv3x3(block_in, 2 * z_channels if double_z else z_channels)

    def forward(self, x_nhwc: torch.Tensor) -> torch.Tensor:
        h = self.conv_in(x_nhwc)
        for i_level in range(self.num_resolutions):
            for blk in self.down[i_level].block:
                h = blk(h)
            if i_level != self.num_resolutions - 1:
                h = self.down[i_level].downsample(h)
        h = self.mid.block_2(self.mid.attn_1(self.mid.block_1(h)))
        return self.conv_out(self.norm_out(h, act=ops.ACT_SILU), out_dtype=torch.float32)


class DiagonalGaussianDistribution:
    """terediff/model/distributions.py:24-60 (mode / sample of the latent posterior)."""

    def __init__(self, parameters: torch.Tensor):
        self.parameters = parameters
        self.mean, logvar = torch.chunk(parameters, 2, dim=1)
        self.logvar = torch.clamp(logvar, -30.0, 20.0)
        self.std = torch.exp(0.5 * self.logvar)

    def mode(self) -> torch.Tensor:
        return self.mean

    def sample(self) -> torch.Tensor:
        return self.mean + self.std * torch.randn_like(self.mean)


class AutoencoderKL(nn.Module):
    """vae.py:562-582 (decode path)."""

    def __init__(self, ddconfig: dict, embed_dim: int):
        super().__init__()
        self.encoder = Encoder(**ddconfig)
        self.decoder = Decoder(**ddconfig)
        assert ddconfig["double_z"]
        self.quant_conv = Conv1x1(2 * ddconfig["z_channels"], 2 * embed_dim)
        self.post_quant_conv = Conv1x1(embed_dim, ddconfig["z_channels"])
        self.embed_dim = embed_dim
        self.z_channels = ddconfig["z_channels"]

    def _post_quant(self):
        """4 -> 4 channel 1x1 conv as a GEMM on the 64-channel padded latent."""
        st = self.post_quant_conv._stamp()
        if getattr(self, "_pq_stamp", None) != st:
            with torch.no_grad():
                w = self.post_quant_conv.weight.detach().reshape(self.z_channels, self.embed_dim)
                wp = torch.zeros((64, 64), device=w.device, dtype=BF16)
                wp[:self.z_channels, :self.embed_dim] = w.to(BF16)
                bp = torch.zeros((64,), device=w.device)
                bp[:self.z_channels] = self.post_quant_conv.bias.detach().float()
                self._pq = (wp.contiguous(), bp.contiguous())
            self._pq_stamp = st
        return self._pq

    @torch.no_grad()
    def encode(self, x: torch.Tensor) -> DiagonalGaussianDistribution:
        """(B,3,H,W) fp32 image in [-1,1] -> posterior over the (B,4,H/8,W/8) latent, as AutoencoderKL.encode."""
        B, _, H, W = x.shape
        h = self.encoder(ops.nchw_to_nhwc(x.float(), 64))                       # [B,H/8,W/8,8] fp32
        hh, ww, cz = h.shape[1], h.shape[2], h.shape[3]
        # quant_conv is an 8 -> 8 channel 1x1 conv on a tiny tensor: fp32 matmul on the moments (torch; 64 FLOP/pixel)
        wq = self.quant_conv.weight.detach().reshape(self.quant_conv.out_channels, cz).float()
        m = h.reshape(-1, cz) @ wq.t() + self.quant_conv.bias.detach().float()
        return DiagonalGaussianDistribution(m.view(B, hh, ww, -1).permute(0, 3, 1, 2).contiguous())

    @torch.no_grad()
    def decode(self, z: torch.Tensor) -> torch.Tensor:
        """(B,4,h,w) fp32 latent -> (B,3,8h,8w) fp32 image, as AutoencoderKL.decode."""
        B, _, h, w = z.shape
        zp = ops.nchw_to_nhwc(z.float(), 64)                                   # [B,h,w,64], channels >= 4 are zero
        wp, bp = self._post_quant()
        zq = ops.gemm(zp.view(B * h * w, 64), wp, bias=bp).view(B, h, w, 64)    # padded channels stay zero
        img = self.decoder(zq)                                                 # [B,8h,8w,3] bf16
        return ops.nhwc_to_nchw(img.contiguous(), 3)

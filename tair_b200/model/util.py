"""Parameter containers with kernel-layout weight packing.

The reference keeps fp32 ``nn.Conv2d`` / ``nn.Linear`` / ``nn.GroupNorm`` / ``nn.LayerNorm`` modules
(terediff/model/util.py:182-216).  Here the same classes are used purely as *parameter containers* so that
``state_dict`` keys and shapes are identical (real checkpoints drop in); their torch ``forward`` is never called.
Each container lazily produces the bf16, K-contiguous layout the sm_100a kernels consume and re-packs when the
underlying parameter is modified (``Tensor._version``) or moved.
"""
from __future__ import annotations

from typing import Optional

import torch
from torch import nn

from .. import ops

BF16 = torch.bfloat16


def _round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class _PackMixin:
    def _stamp(self):
        w = self.weight
        b = getattr(self, "bias", None)
        return (w.data_ptr(), w._version, w.device, None if b is None else (b.data_ptr(), b._version))

    def packed(self):
        st = self._stamp()
        if getattr(self, "_pk_stamp", None) != st:
            with torch.no_grad():
                self._pk = self._pack()
            self._pk_stamp = st
        return self._pk


class Linear(nn.Linear, _PackMixin):
    """nn.Linear container -> (w bf16 [N,K], bias fp32 [N] | None)."""

    def _pack(self):
        w = self.weight.detach().to(BF16).contiguous()
        b = None if self.bias is None else self.bias.detach().float().contiguous()
        return w, b

    def forward(self, x2d: torch.Tensor, *, residual=None, act=ops.ACT_NONE, out=None, out_dtype=BF16):
        w, b = self.packed()
        return ops.gemm(x2d, w, bias=b, residual=residual, act=act, out=out, out_dtype=out_dtype)


class Conv1x1(nn.Conv2d, _PackMixin):
    """1x1 nn.Conv2d container (ResBlock skip, ControlNet zero convs) -> GEMM on [B*H*W, Cin]."""

    def __init__(self, cin, cout):
        super().__init__(cin, cout, 1)

    def _pack(self):
        w = self.weight.detach().reshape(self.out_channels, self.in_channels).to(BF16).contiguous()
        return w, self.bias.detach().float().contiguous()

    def forward(self, x: torch.Tensor, *, residual=None, out=None):
        w, b = self.packed()
        B, H, W, C = x.shape
        o = ops.gemm(x.view(-1, C), w, bias=b, residual=None if residual is None else residual.view(B * H * W, -1),
                     out=None if out is None else out.view(B * H * W, -1))
        return o.view(B, H, W, -1)


class Conv3x3(nn.Conv2d, _PackMixin):
    """3x3 / pad 1 nn.Conv2d container -> implicit-GEMM weight [Cout, 9*Cin_pad], K ordered (ky,kx,ci).

    Cin is zero-padded to a multiple of 64 (the TMA/UMMA k-block); the 4- and 8-channel input convs of the
    UNet / ControlNet (unet.py:491-497, controlnet.py:168-175) therefore run on a 64-channel padded latent.
    """

    def __init__(self, cin, cout, stride=1, asymmetric_pad=False):
        # asymmetric_pad: nn.Conv2d(padding=0) applied after F.pad(x, (0,1,0,1)) — the VAE downsampler (vae.py:40-55)
        super().__init__(cin, cout, 3, stride=stride, padding=0 if asymmetric_pad else 1)
        self.cin_pad = _round_up(cin, 64)
        self.pad_lo = 0 if asymmetric_pad else 1

    def _pack(self):
        w = self.weight.detach()
        co, ci = w.shape[0], w.shape[1]
        wp = torch.zeros((co, 3, 3, self.cin_pad), device=w.device, dtype=BF16)
        wp[..., :ci] = w.permute(0, 2, 3, 1).to(BF16)
        return wp.reshape(co, 9 * self.cin_pad).contiguous(), self.bias.detach().float().contiguous()

    def forward(self, x: torch.Tensor, *, residual=None, rowgroup=None, rows_per_group=0, out=None, out_dtype=BF16):
        w, b = self.packed()
        return ops.conv3x3(x, w, stride=self.stride[0], pad=self.pad_lo, bias=b, residual=residual, rowgroup=rowgroup,
                           rows_per_group=rows_per_group, out=out, out_dtype=out_dtype, real_cin=self.in_channels)


class GroupNorm(nn.GroupNorm, _PackMixin):
    """GroupNorm32 (terediff/model/util.py:182-193, eps 1e-5) / Normalize (attention.py:48-51, eps 1e-6)."""

    def _pack(self):
        return self.weight.detach().float().contiguous(), self.bias.detach().float().contiguous()

    def forward(self, x: torch.Tensor, act=ops.ACT_NONE):
        g, b = self.packed()
        return ops.groupnorm(x, g, b, groups=self.num_groups, eps=self.eps, act=act)


class LayerNorm(nn.LayerNorm, _PackMixin):
    def _pack(self):
        return self.weight.detach().float().contiguous(), self.bias.detach().float().contiguous()

    def forward(self, x2d: torch.Tensor):
        g, b = self.packed()
        return ops.layernorm(x2d, g, b, eps=self.eps)


def normalization(channels: int) -> GroupNorm:
    """terediff/model/util.py:182-189."""
    return GroupNorm(32, channels, eps=1e-5)


def zero_module(m: nn.Module) -> nn.Module:
    """terediff/model/util.py:151-157 — kept so default construction matches the reference initialisation."""
    for p in m.parameters():
        p.detach().zero_()
    return m


def timestep_embedding(timesteps: torch.Tensor, dim: int, max_period: float = 10000.0) -> torch.Tensor:
    """terediff/model/util.py:128-148 -> bf16 [B, dim] on the device."""
    return ops.timestep_embedding(timesteps.to(torch.int64).contiguous(), dim, max_period)

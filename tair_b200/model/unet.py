"""SD2.1 UNet building blocks and trunk on the sm_100a kernels.

Same module tree (and therefore the same ``state_dict`` keys) as terediff/model/unet.py: ``TimestepEmbedSequential``
(:34-48), ``Upsample`` (:51-79), ``Downsample`` (:82-108), ``ResBlock`` (:111-223), ``UNetModel`` (:391-685).
Activations are channels-last bf16 [B,H,W,C]; a ResBlock is five launches:

    GroupNorm+SiLU -> conv3x3 (+bias +timestep-embedding row add, in the epilogue)
    GroupNorm+SiLU -> conv3x3 (+bias +skip, in the epilogue)         [+ one GEMM when the skip is a 1x1 conv]

All ``emb_layers`` Linear(SiLU(emb)) projections of a network are evaluated as ONE GEMM per step, and so are all
cross-attention context K|V projections (see ``UNetModel._begin_step``).
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch
from torch import nn

from .. import ops
from .attention import SpatialTransformer
from .util import BF16, Conv1x1, Conv3x3, Linear, normalization, timestep_embedding, zero_module


class TimestepBlock(nn.Module):
    pass


class TimestepEmbedSequential(nn.Sequential, TimestepBlock):
    """unet.py:34-48 — routes (emb, context) to the children that take them."""

    def forward(self, x, step):
        for layer in self:
            if isinstance(layer, ResBlock):
                x = layer(x, step)
            elif isinstance(layer, SpatialTransformer):
                x = layer(x, step.ctx_kv_for(layer), step.Lk)
            else:
                x = layer(x)
        return x


class Upsample(nn.Module):
    """unet.py:51-79: nearest x2 then conv3x3."""

    def __init__(self, channels: int, use_conv: bool = True, dims: int = 2, out_channels: Optional[int] = None):
        super().__init__()
        assert use_conv and dims == 2
        self.channels, self.out_channels = channels, out_channels or channels
        self.conv = Conv3x3(channels, self.out_channels)

    def forward(self, x):
        return self.conv(ops.upsample2x(x))


class Downsample(nn.Module):
    """unet.py:82-108: conv3x3 stride 2 (TMA elementStrides, no im2col)."""

    def __init__(self, channels: int, use_conv: bool = True, dims: int = 2, out_channels: Optional[int] = None):
        super().__init__()
        assert use_conv and dims == 2
        self.channels, self.out_channels = channels, out_channels or channels
        self.op = Conv3x3(channels, self.out_channels, stride=2)

    def forward(self, x):
        return self.op(x)


class ResBlock(TimestepBlock):
    """unet.py:111-223 without up/down and without scale-shift norm (the val config uses neither)."""

    def __init__(self, channels: int, emb_channels: int, dropout: float = 0.0, out_channels: Optional[int] = None):
        super().__init__()
        self.channels, self.emb_channels = channels, emb_channels
        self.out_channels = out_channels or channels
        self.in_layers = nn.Sequential(normalization(channels), nn.SiLU(), Conv3x3(channels, self.out_channels))
        self.emb_layers = nn.Sequential(nn.SiLU(), Linear(emb_channels, self.out_channels))
        self.out_layers = nn.Sequential(normalization(self.out_channels), nn.SiLU(), nn.Dropout(p=dropout),
                                        zero_module(Conv3x3(self.out_channels, self.out_channels)))
        self.skip_connection = nn.Identity() if self.out_channels == channels else Conv1x1(channels, self.out_channels)

    def forward(self, x, step):
        B, H, W, _ = x.shape
        h = self.in_layers[0](x, act=ops.ACT_SILU)
        h = self.in_layers[2](h, rowgroup=step.emb_for(self), rows_per_group=H * W)
        h = self.out_layers[0](h, act=ops.ACT_SILU)
        skip = x if isinstance(self.skip_connection, nn.Identity) else self.skip_connection(x)
        return self.out_layers[3](h, residual=skip)


class _Step:
    """Per-forward state shared by the blocks of one network: the stacked timestep-embedding projection
    [B, sum(Cout)] fp32 and the stacked cross-attention context projection [B*Lk, sum(2*inner)] bf16."""

    def __init__(self, emb_all, emb_slices, kv_all, kv_slices, Lk):
        self._emb_all, self._emb_slices = emb_all, emb_slices
        self._kv_all, self._kv_slices = kv_all, kv_slices
        self.Lk = Lk

    def emb_for(self, block: ResBlock) -> torch.Tensor:
        a, b = self._emb_slices[id(block)]
        return self._emb_all[:, a:b]

    def ctx_kv_for(self, st: SpatialTransformer):
        return [self._kv_all[:, a:b] for (a, b) in self._kv_slices[id(st)]]


class UNetModel(nn.Module):
    """unet.py:391-685 restricted to the options of configs/val/val_terediff.yaml:6-20 (spatial transformer with
    linear projections, num_head_channels heads, conv resampling, no class conditioning)."""

    def __init__(self, image_size=32, in_channels=4, model_channels=320, out_channels=4, num_res_blocks=2,
                 attention_resolutions=(4, 2, 1), dropout=0, channel_mult=(1, 2, 4, 4), conv_resample=True, dims=2,
                 num_classes=None, use_checkpoint=False, use_fp16=False, num_heads=-1, num_head_channels=-1,
                 num_heads_upsample=-1, use_scale_shift_norm=False, resblock_updown=False,
                 use_new_attention_order=False, use_spatial_transformer=False, transformer_depth=1, context_dim=None,
                 n_embed=None, legacy=True, disable_self_attentions=None, num_attention_blocks=None,
                 disable_middle_self_attn=False, use_linear_in_transformer=False, _build_decoder=True,
                 _hint_channels=0):
        super().__init__()
        unsupported = dict(dims=dims != 2, num_classes=num_classes is not None, use_scale_shift_norm=use_scale_shift_norm,
                           resblock_updown=resblock_updown, no_spatial_transformer=not use_spatial_transformer,
                           conv_transformer_proj=not use_linear_in_transformer, legacy=legacy,
                           num_head_channels=num_head_channels == -1, no_conv_resample=not conv_resample,
                           disable_self_attentions=disable_self_attentions is not None,
                           num_attention_blocks=num_attention_blocks is not None,
                           disable_middle_self_attn=disable_middle_self_attn, n_embed=n_embed is not None)
        bad = [k for k, v in unsupported.items() if v]
        if bad:
            raise NotImplementedError(f"tair_b200 UNet covers the TeReDiff val config only; unsupported: {bad}")
        if isinstance(context_dim, (list, tuple)):
            context_dim = list(context_dim)[0]
        if isinstance(num_res_blocks, int):
            num_res_blocks = len(channel_mult) * [num_res_blocks]
        self.in_channels, self.model_channels, self.out_channels = in_channels, model_channels, out_channels
        self.num_res_blocks, self.channel_mult = list(num_res_blocks), tuple(channel_mult)
        self.attention_resolutions = tuple(attention_resolutions)
        self.context_dim = context_dim
        self.dtype = torch.float32  # dtype of the public (reference-facing) tensors
        ted = model_channels * 4
        self.time_embed = nn.Sequential(Linear(model_channels, ted), nn.SiLU(), Linear(ted, ted))

        def st(ch):
            return SpatialTransformer(ch, ch // num_head_channels, num_head_channels, depth=transformer_depth,
                                      context_dim=context_dim)

        self.input_blocks = nn.ModuleList(
            [TimestepEmbedSequential(Conv3x3(in_channels + _hint_channels, model_channels))])
        chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(self.num_res_blocks[level]):
                layers = [ResBlock(ch, ted, dropout, out_channels=mult * model_channels)]
                ch = mult * model_channels
                if ds in self.attention_resolutions:
                    layers.append(st(ch))
                self.input_blocks.append(TimestepEmbedSequential(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.input_blocks.append(TimestepEmbedSequential(Downsample(ch, True, out_channels=ch)))
                chans.append(ch)
                ds *= 2
        self.middle_block = TimestepEmbedSequential(ResBlock(ch, ted, dropout), st(ch), ResBlock(ch, ted, dropout))
        self._mid_channels = ch
        self._enc_channels = list(chans)
        if _build_decoder:
            self.output_blocks = nn.ModuleList([])
            for level, mult in list(enumerate(channel_mult))[::-1]:
                for i in range(self.num_res_blocks[level] + 1):
                    ich = chans.pop()
                    layers = [ResBlock(ch + ich, ted, dropout, out_channels=model_channels * mult)]
                    ch = model_channels * mult
                    if ds in self.attention_resolutions:
                        layers.append(st(ch))
                    if level and i == self.num_res_blocks[level]:
                        layers.append(Upsample(ch, True, out_channels=ch))
                        ds //= 2
                    self.output_blocks.append(TimestepEmbedSequential(*layers))
            self.out = nn.Sequential(normalization(ch), nn.SiLU(), zero_module(Conv3x3(model_channels, out_channels)))

    # ---- per-step hoisted GEMMs ------------------------------------------------------------------
    def _stacked(self):
        """(emb weight [sum Cout, ted], emb bias, slices) and (ctx K|V weight [sum 2*inner, ctx], slices)."""
        res = [m for m in self.modules() if isinstance(m, ResBlock)]
        sts = [m for m in self.modules() if isinstance(m, SpatialTransformer)]
        stamp = tuple(m.emb_layers[1]._stamp() for m in res) + tuple(
            b.attn2._stamp() for s in sts for b in s.transformer_blocks)
        if getattr(self, "_stk_stamp", None) != stamp:
            with torch.no_grad():
                ew = torch.cat([m.emb_layers[1].weight for m in res], 0).detach().to(BF16).contiguous()
                eb = torch.cat([m.emb_layers[1].bias for m in res], 0).detach().float().contiguous()
                es, o = {}, 0
                for m in res:
                    es[id(m)] = (o, o + m.out_channels)
                    o += m.out_channels
                kws, ks, o = [], {}, 0
                for s in sts:
                    sl = []
                    for b in s.transformer_blocks:
                        w = b.attn2.kv_weight()
                        kws.append(w)
                        sl.append((o, o + w.shape[0]))
                        o += w.shape[0]
                    ks[id(s)] = sl
                kw = torch.cat(kws, 0).contiguous()
            self._stk = (ew, eb, es, kw, ks)
            self._stk_stamp = stamp
        return self._stk

    def _begin_step(self, timesteps: torch.Tensor, context: torch.Tensor) -> _Step:
        """time_embed (util.py:128-148 + unet.py:476-480) and the two hoisted per-network GEMMs."""
        ew, eb, es, kw, ks = self._stacked()
        t_emb = timestep_embedding(timesteps, self.model_channels)
        e = self.time_embed[0](t_emb, act=ops.ACT_SILU)
        semb = self.time_embed[2](e, act=ops.ACT_SILU)  # every consumer applies SiLU first (unet.py:163-169)
        emb_all = ops.gemm(semb, ew, bias=eb, out_dtype=torch.float32)
        B, Lk, D = context.shape
        ctx = context.to(BF16).reshape(B * Lk, D).contiguous()
        kv_all = ops.gemm(ctx, kw)
        return _Step(emb_all, es, kv_all, ks, Lk)

    def _embed_input(self, x: torch.Tensor) -> torch.Tensor:
        """(B,C,H,W) fp32 -> channels-last bf16 padded to the first conv's 64-channel k-block."""
        return ops.nchw_to_nhwc(x.float(), self.input_blocks[0][0].cin_pad)

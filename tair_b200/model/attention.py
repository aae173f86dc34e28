"""SpatialTransformer / BasicTransformerBlock / cross-attention on the sm_100a kernels.

Mirrors terediff/model/attention.py (SpatialTransformer :277-353, BasicTransformerBlock :219-274,
SDPCrossAttention :168-216, GEGLU/FeedForward :19-45) — same submodule names, hence the same state_dict keys —
but operates on channels-last bf16 token matrices [B*HW, C] and launches:

    GroupNorm(eps 1e-6)            tair_groupnorm_nhwc
    proj_in / to_out / ff / proj_out   tair_gemm_bf16 (bias, residual and GEGLU fused into the epilogue)
    to_q|to_k|to_v of attn1        ONE tair_gemm_bf16 with the three weights stacked ([3C, C])
    to_k|to_v of attn2             hoisted: all blocks' context projections are one GEMM per network per step
    softmax(QK^T/8)V               tair_attention_bf16 (tcgen05 flash attention, head_dim 64)
    LayerNorm                      folded into the GEMM that consumes it: norm1 -> attn1 q|k|v, norm2 -> attn2 to_q,
                                   norm3 -> GEGLU.  The GEMM reads the RAW rows with weights W * gamma and rescales its
                                   accumulator per row in the epilogue, rstd * (acc - mean * sum_k W'[n,k]) + (b + W beta);
                                   only the per-row (mean, rstd) are computed beforehand (tair_row_stats: one 2 B/element
                                   read instead of LayerNorm's read + write).  Mathematically identical to
                                   Linear(LayerNorm(x)); the normalised tensor is never materialised.
"""
from __future__ import annotations

import os
from typing import Optional

import torch
from torch import nn

from .. import ops
from .util import BF16, GroupNorm, LayerNorm, Linear, zero_module


def geglu_tile(n_rows: int) -> int:
    """N tile the C side picks for a GEGLU GEMM with ``n_rows`` weight rows (pick_bn in csrc/gemm_tc.cu)."""
    if n_rows % 256 == 0:
        return 256
    if n_rows % 128 == 0:
        return 128
    raise ValueError(f"GEGLU projection with {n_rows} rows cannot be tiled (needs a multiple of 128)")


def fold_layernorm(w: torch.Tensor, b: Optional[torch.Tensor], ln: "LayerNorm"):
    """Linear(LayerNorm(x)) == rstd * (x @ (W*gamma)^T - mean * colsum) + (b + W @ beta): -> (W*gamma as bf16 [N,K],
    bias' fp32 [N], colsum fp32 [N] taken over the bf16-rounded folded weight so that it cancels the accumulator exactly)."""
    w32 = w.detach().float()
    g, beta = ln.weight.detach().float(), ln.bias.detach().float()
    wf = (w32 * g[None, :]).to(BF16).contiguous()
    bias = w32 @ beta
    if b is not None:
        bias = bias + b.detach().float()
    return wf, bias.contiguous(), wf.float().sum(1).contiguous()


def interleave_geglu(w: torch.Tensor, b: torch.Tensor):
    """Reorder GEGLU.proj rows ([value ; gate], attention.py:25-27) so every BN-row tile holds BN/2 value rows
    followed by the BN/2 matching gate rows — the layout the TAIR_ACT_GEGLU epilogue multiplies in registers."""
    n = w.shape[0]
    half = geglu_tile(n) // 2
    inner = n // 2
    idx = torch.arange(inner, device=w.device).view(-1, half)
    perm = torch.cat([idx, idx + inner], dim=1).reshape(-1)
    return w[perm].contiguous(), b[perm].contiguous()


class GEGLU(nn.Module):
    """attention.py:19-27; forward returns proj(x)[..., :d] * gelu(proj(x)[..., d:]) from one fused GEMM."""

    def __init__(self, dim_in: int, dim_out: int):
        super().__init__()
        self.proj = Linear(dim_in, dim_out * 2)

    def _packed(self):
        st = self.proj._stamp()
        if getattr(self, "_pk_stamp", None) != st:
            with torch.no_grad():
                w, b = interleave_geglu(self.proj.weight.detach().to(BF16), self.proj.bias.detach().float())
            self._pk, self._pk_stamp = (w, b), st
        return self._pk

    def _packed_ln(self, ln):
        st = (self.proj._stamp(), ln._stamp())
        if getattr(self, "_ln_stamp", None) != st:
            with torch.no_grad():
                wf, bias, _ = fold_layernorm(self.proj.weight, self.proj.bias, ln)
                w, b = interleave_geglu(wf, bias)
                self._ln_pk = (w, b, w.float().sum(1).contiguous())
            self._ln_stamp = st
        return self._ln_pk

    def forward(self, x2d: torch.Tensor, ln=None, stats=None) -> torch.Tensor:
        """``ln`` (a LayerNorm module) + ``stats`` = row_stats(x2d): computes GEGLU(ln(x2d)) from the raw rows."""
        if ln is not None:
            w, b, cs = self._packed_ln(ln)
            return ops.gemm(x2d, w, bias=b, act=ops.ACT_GEGLU, ln=(stats, cs))
        w, b = self._packed()
        return ops.gemm(x2d, w, bias=b, act=ops.ACT_GEGLU)


class FeedForward(nn.Module):
    """attention.py:29-45 with glu=True: net = [GEGLU, Dropout, Linear]."""

    def __init__(self, dim: int, mult: int = 4):
        super().__init__()
        inner = dim * mult
        self.net = nn.Sequential(GEGLU(dim, inner), nn.Dropout(0.0), Linear(inner, dim))

    def forward(self, x2d: torch.Tensor, residual: torch.Tensor, ln=None, stats=None) -> torch.Tensor:
        return self.net[2](self.net[0](x2d, ln=ln, stats=stats), residual=residual)


class CrossAttention(nn.Module):
    """SDPCrossAttention (attention.py:168-216): bias-free to_q/to_k/to_v, to_out = [Linear, Dropout]."""

    def __init__(self, query_dim: int, context_dim: Optional[int], heads: int, dim_head: int):
        super().__init__()
        inner = heads * dim_head
        self.heads, self.dim_head, self.inner = heads, dim_head, inner
        self.is_self = context_dim is None
        cdim = query_dim if context_dim is None else context_dim
        self.to_q = Linear(query_dim, inner, bias=False)
        self.to_k = Linear(cdim, inner, bias=False)
        self.to_v = Linear(cdim, inner, bias=False)
        self.to_out = nn.Sequential(Linear(inner, query_dim), nn.Dropout(0.0))

    def _stamp(self):
        return (self.to_q._stamp(), self.to_k._stamp(), self.to_v._stamp())

    def qkv_weight(self) -> torch.Tensor:
        """[3*inner, C] stacked to_q|to_k|to_v (self-attention: one GEMM instead of three)."""
        st = self._stamp()
        if getattr(self, "_qkv_stamp", None) != st:
            with torch.no_grad():
                self._qkv = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], 0).detach().to(BF16).contiguous()
            self._qkv_stamp = st
        return self._qkv

    def kv_weight(self) -> torch.Tensor:
        """[2*inner, context_dim] stacked to_k|to_v (consumed by the per-network context GEMM)."""
        st = self._stamp()
        if getattr(self, "_kv_stamp", None) != st:
            with torch.no_grad():
                self._kv = torch.cat([self.to_k.weight, self.to_v.weight], 0).detach().to(BF16).contiguous()
            self._kv_stamp = st
        return self._kv

    def _folded(self, ln):
        """Query-side projection (q|k|v stacked for self-attention, to_q for cross-attention) with ``ln`` folded in."""
        st = (self._stamp(), ln._stamp())
        if getattr(self, "_fold_stamp", None) != st:
            with torch.no_grad():
                w = torch.cat([self.to_q.weight, self.to_k.weight, self.to_v.weight], 0) if self.is_self else self.to_q.weight
                self._fold = fold_layernorm(w, None, ln)
            self._fold_stamp = st
        return self._fold

    def forward(self, x2d: torch.Tensor, B: int, residual: torch.Tensor, kv: Optional[torch.Tensor] = None,
                Lk: int = 0, ln=None) -> torch.Tensor:
        """x2d [B*L, C]; kv = precomputed [B*Lk, 2*inner] context projection for cross attention.  With ``ln`` (a
        LayerNorm module) x2d holds the RAW rows and the normalisation is folded into the projection GEMM."""
        L = x2d.shape[0] // B
        C = self.inner
        scale = self.dim_head ** -0.5
        lnk = None
        if ln is not None:
            wf, bf, cs = self._folded(ln)
            lnk = (ops.row_stats(x2d, ln.eps), cs)
        if kv is None:
            qkv = ops.gemm(x2d, wf, bias=bf, ln=lnk) if ln is not None else ops.gemm(x2d, self.qkv_weight())
            a = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B=B, H=self.heads, Lq=L, Lk=L,
                              head_dim=self.dim_head, scale=scale)
        else:
            q = ops.gemm(x2d, wf, bias=bf, ln=lnk) if ln is not None else self.to_q(x2d)
            a = ops.attention(q, kv[:, :C], kv[:, C:], B=B, H=self.heads, Lq=L, Lk=Lk, head_dim=self.dim_head,
                              scale=scale)
        return self.to_out[0](a, residual=residual)


class BasicTransformerBlock(nn.Module):
    """attention.py:219-274 (gated_ff=True, no self-attn disabling on this path)."""

    def __init__(self, dim: int, n_heads: int, d_head: int, context_dim: int):
        super().__init__()
        self.attn1 = CrossAttention(dim, None, n_heads, d_head)
        self.ff = FeedForward(dim)
        self.attn2 = CrossAttention(dim, context_dim, n_heads, d_head)
        self.norm1, self.norm2, self.norm3 = LayerNorm(dim), LayerNorm(dim), LayerNorm(dim)
        self.fold_layernorm = os.environ.get("TAIR_FOLD_LN", "1") != "0"   # 0: the separate LayerNorm kernel (A/B probe)

    def forward(self, x2d: torch.Tensor, B: int, ctx_kv: torch.Tensor, Lk: int) -> torch.Tensor:
        if not self.fold_layernorm:
            x2d = self.attn1(self.norm1(x2d), B, residual=x2d)
            x2d = self.attn2(self.norm2(x2d), B, residual=x2d, kv=ctx_kv, Lk=Lk)
            return self.ff(self.norm3(x2d), residual=x2d)
        x2d = self.attn1(x2d, B, residual=x2d, ln=self.norm1)
        x2d = self.attn2(x2d, B, residual=x2d, kv=ctx_kv, Lk=Lk, ln=self.norm2)
        return self.ff(x2d, residual=x2d, ln=self.norm3, stats=ops.row_stats(x2d, self.norm3.eps))


class SpatialTransformer(nn.Module):
    """attention.py:277-353 with use_linear=True; x is channels-last [B,H,W,C] bf16."""

    def __init__(self, in_channels: int, n_heads: int, d_head: int, depth: int = 1, context_dim: int = 1024):
        super().__init__()
        inner = n_heads * d_head
        self.in_channels = in_channels
        self.norm = GroupNorm(32, in_channels, eps=1e-6)
        self.proj_in = Linear(in_channels, inner)
        self.transformer_blocks = nn.ModuleList(
            [BasicTransformerBlock(inner, n_heads, d_head, context_dim) for _ in range(depth)])
        self.proj_out = zero_module(Linear(in_channels, inner))  # (sic) argument order as attention.py:331

    def forward(self, x: torch.Tensor, ctx_kv, Lk: int) -> torch.Tensor:
        """ctx_kv: list (one per transformer block) of [B*Lk, 2*inner] context K|V projections."""
        B, H, W, C = x.shape
        x2d = x.view(-1, C)
        h = self.proj_in(self.norm(x).view(-1, C))
        for blk, kv in zip(self.transformer_blocks, ctx_kv):
            h = blk(h, B, kv, Lk)
        return self.proj_out(h, residual=x2d).view(B, H, W, C)

"""OpenCLIP text encoder on the sm_100a kernels (SURVEY.md §8f "next", rank 2; re-run every denoising step to turn the
recognised text into the next step's cross-attention context, spaced_sampler.py:312-317).

Mirrors ``FrozenOpenCLIPEmbedder`` (terediff/model/clip.py:8-61) over the text half of open_clip's ``CLIP``
(terediff/model/open_clip/model.py, transformer.py:199-254): same parameter names (``model.token_embedding``,
``model.positional_embedding``, ``model.transformer.resblocks.N.{ln_1,attn,ln_2,mlp.c_fc,mlp.c_proj}``,
``model.ln_final``, ``model.text_projection``, ``model.logit_scale``).  Per residual block: LayerNorm -> fused in_proj
GEMM -> causal tcgen05 attention (16 heads x 64) -> out_proj GEMM (+residual) -> LayerNorm -> c_fc GEMM (+GELU) ->
c_proj GEMM (+residual).  ``layer='penultimate'`` stops one block early, then ``ln_final`` (clip.py:37-55).

The BPE tokenizer needs open_clip's vocabulary file, which is not shipped here: ``attach_tokenizer`` plugs in any
callable ``List[str] -> LongTensor[B,77]`` (e.g. ``open_clip.tokenize``); ``forward(tokens)`` takes token ids directly.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Callable, List, Optional

import torch
from torch import nn

from .. import ops
from .util import BF16, LayerNorm, Linear


class _MHA(nn.Module):
    """nn.MultiheadAttention parameter container (in_proj_weight / in_proj_bias / out_proj)."""

    def __init__(self, d_model: int, n_head: int):
        super().__init__()
        self.embed_dim, self.num_heads = d_model, n_head
        self.in_proj_weight = nn.Parameter(torch.randn(3 * d_model, d_model) * d_model ** -0.5)
        self.in_proj_bias = nn.Parameter(torch.zeros(3 * d_model))
        self.out_proj = Linear(d_model, d_model)

    def packed(self):
        st = (self.in_proj_weight.data_ptr(), self.in_proj_weight._version, self.in_proj_bias._version)
        if getattr(self, "_pk_stamp", None) != st:
            with torch.no_grad():
                self._pk = (self.in_proj_weight.detach().to(BF16).contiguous(),
                            self.in_proj_bias.detach().float().contiguous())
            self._pk_stamp = st
        return self._pk


class ResidualAttentionBlock(nn.Module):
    """transformer.py:199-254 (no layer scale, self-attention)."""

    def __init__(self, d_model: int, n_head: int, mlp_ratio: float = 4.0):
        super().__init__()
        self.ln_1 = LayerNorm(d_model)
        self.attn = _MHA(d_model, n_head)
        self.ln_2 = LayerNorm(d_model)
        width = int(d_model * mlp_ratio)
        self.mlp = nn.Sequential(OrderedDict([("c_fc", Linear(d_model, width)), ("gelu", nn.GELU()),
                                              ("c_proj", Linear(width, d_model))]))

    def forward(self, x2d: torch.Tensor, B: int, L: int) -> torch.Tensor:
        E, H = self.attn.embed_dim, self.attn.num_heads
        w, b = self.attn.packed()
        qkv = ops.gemm(self.ln_1(x2d), w, bias=b)
        a = ops.attention(qkv[:, :E], qkv[:, E:2 * E], qkv[:, 2 * E:], B=B, H=H, Lq=L, Lk=L, head_dim=E // H, causal=True)
        x2d = self.attn.out_proj(a, residual=x2d)
        return self.mlp.c_proj(self.mlp.c_fc(self.ln_2(x2d), act=ops.ACT_GELU), residual=x2d)


class _Transformer(nn.Module):
    def __init__(self, width: int, layers: int, heads: int):
        super().__init__()
        self.resblocks = nn.ModuleList([ResidualAttentionBlock(width, heads) for _ in range(layers)])


class _TextModel(nn.Module):
    def __init__(self, embed_dim: int, text_cfg: dict):
        super().__init__()
        width, ctx = text_cfg["width"], text_cfg["context_length"]
        if width // text_cfg["heads"] != 64:
            raise NotImplementedError("tair_b200 CLIP text encoder needs 64-wide heads (width 1024, 16 heads)")
        self.context_length, self.vocab_size = ctx, text_cfg["vocab_size"]
        self.positional_embedding = nn.Parameter(torch.randn(ctx, width) * 0.01)
        self.text_projection = nn.Parameter(torch.randn(width, embed_dim) * width ** -0.5)
        self.logit_scale = nn.Parameter(torch.ones([]) * 2.6592)
        self.transformer = _Transformer(width, text_cfg["layers"], text_cfg["heads"])
        self.token_embedding = nn.Embedding(text_cfg["vocab_size"], width)
        self.ln_final = LayerNorm(width)


class FrozenOpenCLIPEmbedder(nn.Module):
    LAYERS = ["last", "penultimate"]

    def __init__(self, embed_dim, vision_cfg=None, text_cfg=None, layer="last"):
        super().__init__()
        assert layer in self.LAYERS
        self.model = _TextModel(embed_dim, dict(text_cfg))
        self.layer = layer
        self.layer_idx = 0 if layer == "last" else 1
        self._tokenizer: Optional[Callable] = None
        # memoised prompt embeddings: prompt -> row of ONE preallocated [cache_size, 77, width] fp32 buffer.  Rows are
        # written in place, so a sampling run allocates nothing per step (keeping every step's embeddings as separate
        # tensors grew the allocator by 5 MB per step at 16 tiles, and the cudaMalloc calls that followed cost 3 - 80 ms
        # each inside the loop)
        self._cache: "OrderedDict[str, int]" = OrderedDict()
        self._rows: Optional[torch.Tensor] = None
        self._free: List[int] = []
        self._cache_stamp = None
        self.cache_size = 256
        self._plist = None
        self.prompts_requested = 0      # counters (bench / tools): prompts asked for and prompts actually run
        self.prompts_encoded = 0
        # ~190 launches of mostly tiny kernels: launch-bound when run eagerly (3.8 ms for one prompt vs 0.4 ms replayed),
        # so the transformer is captured once per batch size and replayed
        self.use_cuda_graph = True
        self._graphs: "OrderedDict[tuple, tuple]" = OrderedDict()

    def attach_tokenizer(self, fn: Callable[[List[str]], torch.Tensor]) -> None:
        self._tokenizer = fn
        self.clear_cache()

    def clear_cache(self) -> None:
        """Drop memoised prompt embeddings (done automatically when the weights change)."""
        self._cache.clear()
        self._free = list(range(self._rows.shape[0])) if self._rows is not None else []

    def _reserve(self, n_rows: int) -> None:
        """Make the row buffer hold at least ``n_rows`` prompts on the weights' device (re-allocation drops the memo)."""
        pe = self.model.positional_embedding          # [context length, width]: the shape of one prompt's embedding
        cap = max(self.cache_size, n_rows)
        if self._rows is None or self._rows.device != pe.device or self._rows.shape[0] < cap \
                or self._rows.shape[1:] != pe.shape:
            self._rows = torch.empty((cap,) + tuple(pe.shape), device=pe.device, dtype=torch.float32)
            self.clear_cache()

    def _weight_stamp(self) -> tuple:
        """Changes whenever any parameter is modified in place (``_version``), re-allocated (``data_ptr``: load into a new
        module, ``.to()``) or cast: keys both the captured graphs and the memoised prompt embeddings.  Runs once per
        ``encode`` on the sampler's critical path, hence the cached parameter list (walking the module tree costs 0.9 ms,
        the stamp over the cached list 0.1 ms); ``_apply`` / ``load_state_dict`` drop that list."""
        if self._plist is None:
            self._plist = list(self.model.parameters())
        return tuple([(p.data_ptr(), p._version, p.dtype) for p in self._plist])

    def _apply(self, fn, *args, **kwargs):
        self._plist = None
        return super()._apply(fn, *args, **kwargs)

    def load_state_dict(self, *args, **kwargs):
        self._plist = None
        return super().load_state_dict(*args, **kwargs)

    # new prompts per step vary between 1 and the tile batch: pad the batch to a few bucket sizes so that at most
    # len(_BUCKETS) graphs exist per weight stamp instead of one per distinct count
    _BUCKETS = (1, 2, 4, 8, 16, 32, 64)

    @torch.no_grad()
    def forward(self, tokens: torch.Tensor, borrow: bool = False, stamp: Optional[tuple] = None) -> torch.Tensor:
        """``borrow``: return a view of the graph's static output (valid until the next call) instead of a copy;
        ``stamp``: the caller's fresh ``_weight_stamp()`` (saves computing it twice per ``encode``)."""
        if not (self.use_cuda_graph and tokens.is_cuda) or torch.cuda.is_current_stream_capturing():
            return self.encode_with_transformer(tokens)
        n = tokens.shape[0]
        nb = next((b for b in self._BUCKETS if b >= n), n)
        key = (nb, tokens.shape[1], tokens.device, stamp if stamp is not None else self._weight_stamp())
        entry = self._graphs.pop(key, None)
        if entry is None:
            buf = tokens.new_zeros((nb, tokens.shape[1]))
            side = self._side = getattr(self, "_side", None) or torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):   # warm-up outside capture: weight packing, tile tuning, allocator
                for _ in range(2):
                    self.encode_with_transformer(buf)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, stream=side):   # same stream as the warm-up: its workspaces exist outside the capture
                out = self.encode_with_transformer(buf)
            while len(self._graphs) >= 8:                       # least recently used first
                self._graphs.pop(next(iter(self._graphs)))
            entry = (g, buf, out)
        self._graphs[key] = entry                               # (re-)insert at the most-recently-used end
        g, buf, out = entry
        buf[:n].copy_(tokens)
        g.replay()
        return out[:n] if borrow else out[:n].clone()

    @torch.no_grad()
    def encode_with_transformer(self, text: torch.Tensor) -> torch.Tensor:
        """tokens (B,77) int64 -> (B,77,width) fp32 — clip.py:37-45."""
        m = self.model
        B, L = text.shape
        x = (m.token_embedding(text) + m.positional_embedding).to(BF16).reshape(B * L, -1).contiguous()
        blocks = m.transformer.resblocks
        for i, blk in enumerate(blocks):
            if i == len(blocks) - self.layer_idx:
                break
            x = blk(x, B, L)
        return m.ln_final(x).float().view(B, L, -1)

    def encode(self, text: List[str]) -> torch.Tensor:
        """clip.py:56-61, batched: the sampler re-encodes one prompt per tile per step and most of them repeat from
        step to step, so embeddings are memoised by prompt string and only the distinct new strings are run (as one
        batch) through the transformer."""
        if self._tokenizer is None:
            import os
            if os.environ.get("TAIR_BPE_VOCAB"):
                from ..tokenizer import BPETokenizer
                self._tokenizer = BPETokenizer()
            else:
                raise RuntimeError("FrozenOpenCLIPEmbedder.encode needs a BPE tokenizer: set TAIR_BPE_VOCAB to the CLIP "
                                   "merge table or call attach_tokenizer(fn) with a List[str] -> LongTensor[B,77] callable")
        if isinstance(text, str):
            text = [text]
        stamp = self._weight_stamp()
        if stamp != self._cache_stamp:      # weights changed (load_state_dict / .to() / cast): drop stale embeddings
            self.clear_cache()
            self._cache_stamp = stamp
        distinct = list(dict.fromkeys(text))
        self._reserve(len(distinct))
        self.prompts_requested += len(text)
        for t in distinct:                  # this call's prompts become the most recently used: never evicted below
            if t in self._cache:
                self._cache.move_to_end(t)
        new = [t for t in distinct if t not in self._cache]
        while len(self._free) < len(new):   # least recently used first
            self._free.append(self._cache.popitem(last=False)[1])
        for t in new:
            self._cache[t] = self._free.pop()
        # one small H2D carries both index lists; rows are copied in and gathered out on the device
        idx = torch.tensor([self._cache[t] for t in new] + [self._cache[t] for t in text], dtype=torch.long,
                           device=self._rows.device)
        if new:
            self.prompts_encoded += len(new)
            z = self(self._tokenizer(new).to(self._rows.device), borrow=True, stamp=stamp)
            self._rows.index_copy_(0, idx[:len(new)], z)
        return self._rows.index_select(0, idx[len(new):])

"""OpenCLIP text encoder on the sm_100a kernels (tair_b200.model.clip) against the oracle and the reference fixture,
plus the causal mode of the tcgen05 attention kernel it relies on."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TEXT_CFG = dict(context_length=77, vocab_size=49408, width=1024, heads=16, layers=24)
VISION_CFG = dict(image_size=224, layers=32, width=1280, head_width=80, patch_size=14)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


@pytest.mark.parametrize("B,H,L", [(2, 16, 77), (1, 5, 128), (3, 2, 300), (1, 1, 1), (2, 3, 129)])
def test_causal_attention(cuda_lib, B, H, L):
    from tair_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(L)
    C = H * 64
    qkv = (torch.randn(B * L, 3 * C, device="cuda", generator=g) * 1.3).bfloat16()
    out = ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B=B, H=H, Lq=L, Lk=L, causal=True)

    def heads(t):
        return t.float().reshape(B, L, H, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(heads(qkv[:, :C]), heads(qkv[:, C:2 * C]), heads(qkv[:, 2 * C:]), is_causal=True)
    assert torch.isfinite(out.float()).all()
    assert rel(out, ref.transpose(1, 2).reshape(B * L, C)) < 2e-2


def test_causal_needs_square(cuda_lib):
    from tair_b200 import ops
    q = torch.zeros(64, 64, device="cuda", dtype=torch.bfloat16)
    k = torch.zeros(128, 64, device="cuda", dtype=torch.bfloat16)
    with pytest.raises(RuntimeError):
        ops.attention(q, k, k, B=1, H=1, Lq=64, Lk=128, causal=True)


@pytest.fixture(scope="module")
def clip(cuda_lib, manifests):
    from oracle import weights
    from tair_b200.model.clip import FrozenOpenCLIPEmbedder
    sd = weights.seeded_state_dict(manifests["clip_text"])
    m = FrozenOpenCLIPEmbedder(1024, VISION_CFG, TEXT_CFG, layer="penultimate")
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == manifests["clip_text"]
    m.load_state_dict(sd)
    return m.cuda().eval(), {k: v.cuda() for k, v in sd.items()}


def test_clip_vs_reference_fixture_and_oracle(clip, golden):
    from oracle import clip as OC
    m, sd = clip
    g = golden("clip_text.npz")
    tokens = torch.from_numpy(g["tokens"]).cuda()
    z = m(tokens)
    assert z.shape == (2, 77, 1024) and z.dtype == torch.float32
    assert rel(z.cpu()[:, ::4, ::2], torch.from_numpy(g["z"])) < 4e-2
    with torch.no_grad():
        ref = OC.encode_tokens(sd, tokens)
    assert rel(z, ref) < 4e-2
    # causal: a token's embedding must not depend on later tokens
    t2 = tokens.clone()
    t2[:, 40:] = 5
    z2 = m(t2)
    assert torch.equal(z2[:, :40], z[:, :40])


def test_clip_last_layer_and_tokenizer_hook(clip):
    from oracle import clip as OC
    from tair_b200.model.clip import FrozenOpenCLIPEmbedder
    m, sd = clip
    last = FrozenOpenCLIPEmbedder(1024, VISION_CFG, TEXT_CFG, layer="last").cuda().eval()
    last.load_state_dict(m.state_dict())
    tokens = torch.randint(0, 49408, (16, 77), device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    with torch.no_grad():
        ref = OC.encode_tokens(sd, tokens, layer="last")
    assert rel(last(tokens), ref) < 4e-2
    with pytest.raises(RuntimeError):
        m.encode(["a sign"])
    m.attach_tokenizer(lambda texts: tokens[:len(texts)].cpu())
    assert torch.equal(m.encode(["x", "y"]), m(tokens[:2]))
    m.attach_tokenizer(None)


def test_clip_prompt_memo_rows_eviction_and_weight_change(clip):
    """encode() memoises prompt embeddings in ONE preallocated row buffer: repeated and re-ordered prompts come back
    identical, the least recently used rows are evicted first (never a prompt of the current call), a call with more
    distinct prompts than rows grows the buffer, and changing a weight in place drops every memoised row."""
    import zlib
    from tair_b200.model.clip import FrozenOpenCLIPEmbedder
    m, _ = clip
    e = FrozenOpenCLIPEmbedder(1024, VISION_CFG, TEXT_CFG, layer="penultimate").cuda().eval()
    e.load_state_dict(m.state_dict())
    e.cache_size = 6

    def tok(texts):
        out = torch.zeros((len(texts), 77), dtype=torch.long)
        for i, t in enumerate(texts):
            g = torch.Generator().manual_seed(zlib.crc32(t.encode()))
            out[i, :8] = torch.randint(1, 49000, (8,), generator=g)
        return out
    calls = []
    e.attach_tokenizer(lambda texts: (calls.append(list(texts)), tok(texts))[1])
    direct = lambda texts: e(tok(texts).cuda())                         # noqa: E731
    a = e.encode(["p0", "p1", "p2", "p1"])
    assert calls == [["p0", "p1", "p2"]] and a.shape == (4, 77, 1024)
    assert torch.equal(a, direct(["p0", "p1", "p2", "p1"]))
    ptr = e._rows.data_ptr()
    b = e.encode(["p2", "p0", "p3"])                                    # two hits, one new
    assert calls[-1] == ["p3"] and torch.equal(b, direct(["p2", "p0", "p3"]))
    e.encode(["p4", "p5", "p6"])                                        # 7 prompts, 6 rows: p1 (least recent) leaves
    assert "p1" not in e._cache and len(e._cache) == 6 and e._rows.data_ptr() == ptr
    c = e.encode(["p0", "p1"])                                          # p0 still memoised; p1 re-encoded, p2 leaves
    assert calls[-1] == ["p1"] and "p2" not in e._cache
    assert torch.equal(c, direct(["p0", "p1"]))
    many = [f"q{i}" for i in range(9)]                                  # more distinct prompts than rows: buffer grows
    d = e.encode(many)
    assert e._rows.shape[0] == 9 and torch.equal(d, direct(many))
    n_calls = len(calls)
    assert torch.equal(e.encode(many[::-1]), d.flip(0)) and len(calls) == n_calls
    with torch.no_grad():
        e.model.ln_final.weight.mul_(2.0)                               # in-place weight change: memo must be dropped
    d2 = e.encode(many)
    assert len(calls) == n_calls + 1 and not torch.equal(d2, d) and torch.equal(d2, direct(many))

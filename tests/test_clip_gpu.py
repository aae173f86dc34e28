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

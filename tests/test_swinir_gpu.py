"""SwinIR cleaner on the sm_100a kernels (tair_b200.model.swinir) against the oracle and the reference fixture, plus
the kernels it adds: window attention with a bias table, ragged LayerNorm, row gather, LeakyReLU."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

CFG = dict(img_size=64, patch_size=1, in_chans=3, embed_dim=180, depths=[6] * 8, num_heads=[6] * 8, window_size=8,
           mlp_ratio=2, sf=8, img_range=1.0, upsampler="nearest+conv", resi_connection="1conv", unshuffle=True,
           unshuffle_scale=8)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def test_layernorm_ragged_gather_leaky(cuda_lib):
    from tair_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.zeros(1000, 192, device="cuda")
    x[:, :180] = torch.randn(1000, 180, device="cuda", generator=g) * 2 + 0.5
    gam, bet = torch.zeros(192, device="cuda"), torch.zeros(192, device="cuda")
    gam[:180] = 1 + 0.1 * torch.randn(180, device="cuda", generator=g)
    bet[:180] = 0.1 * torch.randn(180, device="cuda", generator=g)
    y = ops.layernorm_ragged(x.bfloat16(), gam, bet, 180)
    ref = F.layer_norm(x.bfloat16().float()[:, :180], (180,), gam[:180], bet[:180], 1e-5)
    assert rel(y[:, :180], ref) < 1e-2 and y[:, 180:].abs().max() == 0
    idx = torch.randperm(1000, device="cuda", generator=g).to(torch.int32)
    xb = x.bfloat16()
    assert torch.equal(ops.gather_rows(xb, idx), xb[idx.long()])
    assert torch.equal(ops.leaky_relu(xb, 0.2), F.leaky_relu(xb.float(), 0.2).bfloat16())


@pytest.mark.parametrize("nwin,nw_tbl", [(64, 1), (130, 64), (3, 2)])
def test_attention_windows_with_bias_table(cuda_lib, nwin, nw_tbl):
    from tair_b200 import ops
    H, L, hd = 6, 64, 30
    g = torch.Generator(device="cuda").manual_seed(nwin)
    real = torch.randn(nwin * L, 3, H, hd, device="cuda", generator=g).bfloat16()
    qkv = torch.zeros(nwin * L, 3, H, 64, device="cuda", dtype=torch.bfloat16)
    qkv[..., :hd] = real
    bias = torch.randn(nw_tbl, H, L, L, device="cuda", generator=g)          # [table, head, query, key], natural units
    bias[:, :, :, ::7] -= 100.0 * (torch.rand(nw_tbl, H, L, (L + 6) // 7, device="cuda", generator=g) < 0.3)
    scale = hd ** -0.5
    tbl = (bias.transpose(2, 3) / scale).contiguous()                        # kernel layout [table, head, key, query] / scale
    out = ops.attention_windows(qkv.view(nwin * L, -1), n_heads=H, L=L, n_windows=nwin, scale=scale, bias=tbl)
    q, k, v = (real[:, i].float().view(nwin, L, H, hd).transpose(1, 2) for i in range(3))
    att = q @ k.transpose(-1, -2) * scale + bias[torch.arange(nwin, device="cuda") % nw_tbl]
    ref = (torch.softmax(att, -1) @ v).transpose(1, 2).reshape(nwin * L, H, hd)
    got = out.view(nwin * L, H, 64)
    assert got[..., hd:].abs().max() == 0
    assert rel(got[..., :hd], ref) < 1.5e-2


@pytest.fixture(scope="module")
def swinir(cuda_lib, manifests):
    from oracle import weights
    from tair_b200.model.swinir import SwinIR
    sd = weights.seeded_state_dict(manifests["swinir"])
    m = SwinIR(**CFG)
    assert {k: list(v.shape) for k, v in m.state_dict().items() if not k.endswith("relative_position_index")} == manifests["swinir"]
    own = m.state_dict()
    for k in sd:   # the 0/-100 shift masks are constants of the architecture, keep ours
        if k.endswith("attn_mask"):
            sd[k] = own[k]
    missing, unexpected = m.load_state_dict(sd, strict=False)
    assert not unexpected and all(k.endswith("relative_position_index") for k in missing)
    return m.cuda().eval(), {k: v.cuda() for k, v in sd.items()}


def test_swinir_vs_reference_fixture_and_oracle(swinir, golden):
    from oracle import swinir as OS
    m, sd = swinir
    g = golden("swinir.npz")
    x = torch.from_numpy(g["x"]).cuda()
    y = m(x)
    assert y.shape == (1, 3, 128, 128) and y.dtype == torch.float32
    assert rel(y.cpu(), torch.from_numpy(g["y"])) < 4e-2
    with torch.no_grad():
        ref = OS.swinir_forward(sd, x)
    assert rel(y, ref) < 4e-2


def test_swinir_full_tiles_batch(swinir):
    """Two 512x512 tiles (the per-tile cleaner call of val_patches.py:324): PSNR against the fp32 oracle, and batch
    independence of the result."""
    from oracle import swinir as OS, vae as OV
    m, sd = swinir
    x = torch.rand((2, 3, 512, 512), device="cuda", generator=torch.Generator(device="cuda").manual_seed(9))
    y = m(x)
    with torch.no_grad():
        ref = OS.swinir_forward(sd, x)
    p = OV.psnr(y.clamp(0, 1), ref.clamp(0, 1)).min().item()
    print(f"SwinIR 512^2 PSNR vs fp32 oracle: {p:.1f} dB, rel max-abs {rel(y, ref):.2e}")
    assert p >= 40.0
    assert torch.equal(m(x[1:])[0], y[1])
    with pytest.raises(ValueError):
        m(x[:, :, :100, :100])

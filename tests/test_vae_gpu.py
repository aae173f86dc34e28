"""VAE decoder (tair_b200.model.vae, sm_100a kernels) against the oracle restatement and the reference fixture."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

DD = dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
          num_res_blocks=2, attn_resolutions=[], dropout=0.0)


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


@pytest.fixture(scope="module")
def vae(cuda_lib, manifests):
    from oracle import weights
    from tair_b200.model.vae import AutoencoderKL
    sd = weights.seeded_state_dict(manifests["vae"])
    m = AutoencoderKL(DD, 4)
    m.load_state_dict(sd)
    assert {k: list(v.shape) for k, v in m.state_dict().items()} == manifests["vae"]
    return m.cuda().eval(), {k: v.cuda() for k, v in sd.items()}


def test_softmax_rows_and_transpose(cuda_lib):
    from tair_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = (torch.randn(300, 4096, device="cuda", generator=g) * 3).bfloat16()
    ref = torch.softmax(x.float() * 0.044, dim=-1)
    assert rel(ops.softmax_rows(x, scale=0.044), ref) < 1e-2
    y = torch.randn(3, 130, 72, device="cuda", generator=g).bfloat16()
    assert torch.equal(ops.transpose(y), y.transpose(1, 2).contiguous())


def test_decode_small_vs_reference_fixture_and_oracle(vae, golden):
    from oracle import vae as OV, weights
    m, sd = vae
    g = golden("vae_decode.npz")
    z = weights.seeded_randn((1, 4, 16, 16), 51).cuda()
    img = m.decode(z)
    assert img.shape == (1, 3, 128, 128) and img.dtype == torch.float32
    assert rel(img.cpu()[:, :, ::2, ::2], torch.from_numpy(g["img"])) < 4e-2
    with torch.no_grad():
        ref = OV.vae_decode(sd, z)
    assert rel(img, ref) < 4e-2


def test_decode_full_tile_psnr(vae):
    """64x64 latents -> 512x512 images (the per-tile decode of val_patches.py:369), B=2, PSNR against the fp32 oracle."""
    from oracle import vae as OV, weights
    m, sd = vae
    z = weights.seeded_randn((2, 4, 64, 64), 61).cuda() * 0.18215 * 4
    img = ((m.decode(z / 0.18215) + 1) / 2).clamp(0, 1)
    with torch.no_grad():
        ref = OV.latent_to_image(sd, z)
    assert img.shape == (2, 3, 512, 512)
    p = OV.psnr(img, ref).min().item()
    print(f"VAE decode 512x512: PSNR {p:.1f} dB vs the fp32 oracle")
    assert p >= 40.0, p   # the north star's final-image gate applied to the decoder alone


def test_encode_vs_reference_fixture_and_oracle(vae, golden):
    """AutoencoderKL.encode(...).mode(): the c_img branch of prepare_condition (cldm.py:143-158)."""
    from oracle import vae as OV, weights
    m, sd = vae
    g = golden("vae_decode.npz")
    x = weights.seeded_randn((1, 3, 64, 64), 52).clamp(-1, 1).cuda()
    post = m.encode(x)
    assert post.parameters.shape == (1, 8, 8, 8) and post.mode().shape == (1, 4, 8, 8)
    assert rel(post.parameters.cpu(), torch.from_numpy(g["moments"])) < 4e-2
    xb = weights.seeded_randn((2, 3, 256, 256), 53).clamp(-1, 1).cuda()
    with torch.no_grad():
        ref = OV.vae_encode_moments(sd, xb)
    assert rel(m.encode(xb).parameters, ref) < 4e-2


def test_asymmetric_pad_conv(cuda_lib):
    import torch.nn.functional as F
    from tair_b200 import ops
    g = torch.Generator(device="cuda").manual_seed(0)
    x = torch.randn(2, 128, 64, 64, device="cuda", generator=g).bfloat16()
    w = (torch.randn(128, 128, 3, 3, device="cuda", generator=g) / 34).bfloat16()
    out = ops.conv3x3(x.permute(0, 2, 3, 1).contiguous(), w.permute(0, 2, 3, 1).reshape(128, -1).contiguous(), stride=2, pad=0)
    ref = F.conv2d(F.pad(x.float(), (0, 1, 0, 1)), w.float(), stride=2).permute(0, 2, 3, 1)
    assert out.shape == (2, 32, 32, 128) and rel(out, ref) < 1e-2

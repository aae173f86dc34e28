"""End-to-end parity of ControlNet + controlled UNet (tair_b200.model, sm_100a kernels) against the oracle and the
reference-generated fixture, on identical seeded weights and inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

# bf16 activations through ~60 layers: max-abs error relative to max-abs of the fp32 result
UNET_TOL = 4e-2
FEAT_TOL = 6e-2


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def cfgs(mc, ctx):
    u = dict(in_channels=4, out_channels=4, model_channels=mc, attention_resolutions=[4, 2, 1], num_res_blocks=2,
             channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
             use_linear_in_transformer=True, transformer_depth=1, context_dim=ctx, legacy=False)
    c = dict(u)
    c.pop("out_channels")
    c["hint_channels"] = 4
    return u, c


def build(mc, ctx, manifests, tag):
    from oracle import weights
    from tair_b200.model import ControlLDM
    usd = weights.seeded_state_dict(manifests[f"unet_{tag}"])
    csd = weights.seeded_state_dict(manifests[f"controlnet_{tag}"])
    m = ControlLDM(*cfgs(mc, ctx))
    m.unet.load_state_dict(usd)
    m.controlnet.load_state_dict(csd)
    return m.cuda().eval(), usd, csd


def test_state_dict_keys_match_reference_manifest(manifests):
    from tair_b200.model import ControlLDM
    m = ControlLDM(*cfgs(320, 1024))
    assert {k: list(v.shape) for k, v in m.unet.state_dict().items()} == manifests["unet_full"]
    assert {k: list(v.shape) for k, v in m.controlnet.state_dict().items()} == manifests["controlnet_full"]


def test_narrow_config_vs_reference_fixture_and_oracle(cuda_lib, golden, manifests):
    from oracle import unet as O, weights
    g = golden("unet_narrow.npz")
    m, usd, csd = build(64, 128, manifests, "narrow")
    x, hint, ctx = (weights.seeded_randn(s, i).cuda() for s, i in (((2, 4, 32, 32), 1), ((2, 4, 32, 32), 2), ((2, 77, 128), 3)))
    t = torch.from_numpy(g["t"]).cuda()
    eps, feats = m(x, t, dict(c_txt=ctx, c_img=hint))
    assert rel(eps.cpu(), torch.from_numpy(g["out"])) < UNET_TOL
    for i, f in enumerate(feats):
        assert rel(f[:, ::4, ::2, ::2].cpu(), torch.from_numpy(g[f"feat{i}"])) < FEAT_TOL, i
    # drop-in surfaces of the two sub-networks (reference tensor conventions, controlnet.py:18-56,323-337)
    ctrl = m.controlnet(x=x, hint=hint, timesteps=t, context=ctx)
    assert len(ctrl) == 13 and ctrl[0].shape == (2, 64, 32, 32) and ctrl[12].shape == (2, 256, 4, 4)
    assert rel(ctrl[0][:, ::4, ::2, ::2].cpu(), torch.from_numpy(g["ctrl0"])) < UNET_TOL
    assert rel(ctrl[12].cpu()[:, ::4], torch.from_numpy(g["ctrl12"])) < UNET_TOL
    out2, feats2 = m.unet(x=x, timesteps=t, context=ctx, control=[c.clone() for c in ctrl], only_mid_control=False)
    assert rel(out2.cpu(), torch.from_numpy(g["out"])) < UNET_TOL
    out3, _ = m.unet(x=x, timesteps=t, context=ctx, control=None)
    with torch.no_grad():
        ref3, _ = O.unet_forward({k: v.cuda() for k, v in usd.items()}, x, t, ctx, None)
    assert rel(out3, ref3) < UNET_TOL


@pytest.mark.parametrize("B", [1, 2])
def test_full_config_vs_oracle(cuda_lib, manifests, B):
    """configs/val/val_terediff.yaml geometry: 320 ch, 64x64 latent, ctx 77x1024; oracle = fp32 torch on the GPU."""
    from oracle import unet as O, weights
    m, usd, csd = build(320, 1024, manifests, "full")
    usd = {k: v.cuda() for k, v in usd.items()}
    csd = {k: v.cuda() for k, v in csd.items()}
    x, hint, ctx = (weights.seeded_randn(s, i).cuda() for s, i in (((B, 4, 64, 64), 1), ((B, 4, 64, 64), 2), ((B, 77, 1024), 3)))
    t = torch.tensor([979, 20][:B], device="cuda")
    eps, feats = m(x, t, dict(c_txt=ctx, c_img=hint))
    with torch.no_grad():
        ref, rfeats = O.cldm_forward(usd, csd, x, t, ctx, hint)
    assert [tuple(f.shape[1:]) for f in feats] == [(1280, 16, 16), (1280, 32, 32), (640, 64, 64), (320, 64, 64)]
    assert rel(eps, ref) < UNET_TOL
    for f, r in zip(feats, rfeats):
        assert rel(f, r) < FEAT_TOL

"""The legacy DiffBIR-style surface (terediff/pipeline.py:45-397, terediff/utils/common.py:125-234) on the kernels:
``SwinIRPipeline.run`` end to end on a narrow model, untiled and with latent tiling (``make_tiled_fn`` through
``SpacedSampler.sample(tiled=True)``), and the tiled model against the untiled one on a latent a single tile covers."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

VAE = dict(ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128, ch_mult=[1, 2, 4, 4],
                         num_res_blocks=2, attn_resolutions=[], dropout=0.0), embed_dim=4)


class StubClip:
    def encode(self, prompts):
        if isinstance(prompts, str):
            prompts = [prompts]
        g = torch.Generator(device="cuda").manual_seed(len(prompts[0]) + 1)
        return torch.randn((1, 77, 128), device="cuda", generator=g).repeat(len(prompts), 1, 1)


@pytest.fixture(scope="module")
def parts(cuda_lib, manifests):
    from oracle import weights
    from tair_b200.model import ControlLDM
    from tair_b200.model.gaussian_diffusion import val_diffusion
    u = dict(in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[4, 2, 1], num_res_blocks=2,
             channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True, use_linear_in_transformer=True,
             transformer_depth=1, context_dim=128, legacy=False)
    c = dict(u)
    c.pop("out_channels")
    c["hint_channels"] = 4
    m = ControlLDM(u, VAE, None, c)
    m.unet.load_state_dict(weights.seeded_state_dict(manifests["unet_narrow"]))
    m.controlnet.load_state_dict(weights.seeded_state_dict(manifests["controlnet_narrow"]))
    m.vae.load_state_dict(weights.seeded_state_dict(manifests["vae"]))
    m = m.cuda().eval()
    m.attach_clip(StubClip())
    return m, val_diffusion()


ARGS = dict(steps=3, strength=1.0, cleaner_tiled=False, cleaner_tile_size=512, cleaner_tile_stride=256,
            vae_encoder_tiled=False, vae_encoder_tile_size=256, vae_decoder_tiled=False, vae_decoder_tile_size=256,
            cldm_tiled=False, cldm_tile_size=512, cldm_tile_stride=256, pos_prompt="a photo", neg_prompt="low quality",
            cfg_scale=1.0, start_point_type="noise", sampler_type="spaced", noise_aug=0, rescale_cfg=False)


def test_pipeline_run_untiled_and_latent_tiled(parts):
    from tair_b200.legacy import SwinIRPipeline
    m, diffusion = parts
    cleaner = lambda x: x.clamp(0, 1)       # identity stage-1 model: the pipeline plumbing is what is under test
    pipe = SwinIRPipeline(cleaner, m, diffusion, None, "cuda")
    lq = np.random.default_rng(0).integers(0, 256, (1, 512, 512, 3), dtype=np.uint8)
    torch.manual_seed(0)
    out = pipe.run(lq, **ARGS)
    assert out.shape == (1, 512, 512, 3) and out.dtype == np.uint8
    # classifier-free guidance + start from the noised condition (pipeline.py:148-160) + noise augmentation
    torch.manual_seed(0)
    out2 = pipe.run(lq, **{**ARGS, "cfg_scale": 4.0, "start_point_type": "cond", "noise_aug": 10})
    assert out2.shape == (1, 512, 512, 3)
    # a 512 x 1024 image with latent tiling: 64 x 128 latent -> three 64 x 64 tiles at stride 32, gaussian-weighted
    lq_wide = np.random.default_rng(1).integers(0, 256, (1, 512, 1024, 3), dtype=np.uint8)
    torch.manual_seed(0)
    out3 = pipe.run(lq_wide, **{**ARGS, "cldm_tiled": True})
    assert out3.shape == (1, 512, 1024, 3)
    assert m.control_scales == [1.0] * 13


def test_tiled_model_equals_untiled_on_one_tile(parts):
    """With a tile that covers the whole latent, make_tiled_fn's weighted average is the identity."""
    from tair_b200.legacy import tiled_model
    m, _ = parts
    g = torch.Generator(device="cuda").manual_seed(3)
    x = torch.randn((1, 4, 64, 64), device="cuda", generator=g)
    cond = dict(c_txt=torch.randn((1, 77, 128), device="cuda", generator=g), c_img=torch.randn((1, 4, 64, 64), device="cuda", generator=g))
    t = torch.full((1,), 500, device="cuda", dtype=torch.long)
    ref, _ = m(x, t, cond)
    out, feats = tiled_model(m, 64, 32)(x, t, cond)
    assert feats is None and torch.allclose(out, ref, atol=1e-5, rtol=1e-5)

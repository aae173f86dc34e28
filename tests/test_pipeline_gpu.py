"""Tile driver on the GPU: restore_image must not depend on how tiles are batched (noise keyed by the global tile
index, batch-independent kernels), and merge_patches_with_overlap keeps the reference signature."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def narrow_model(manifests):
    from oracle import weights
    from tair_b200.model import ControlLDM
    u = dict(in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[4, 2, 1], num_res_blocks=2,
             channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
             use_linear_in_transformer=True, transformer_depth=1, context_dim=128, legacy=False)
    c = dict(u)
    c.pop("out_channels")
    c["hint_channels"] = 4
    m = ControlLDM(u, c)
    m.unet.load_state_dict(weights.seeded_state_dict(manifests["unet_narrow"]))
    m.controlnet.load_state_dict(weights.seeded_state_dict(manifests["controlnet_narrow"]))
    return m.cuda().eval()


def cond_fn(x):
    c_img = F.avg_pool2d(x, 8).mean(1, keepdim=True).repeat(1, 4, 1, 1) * 2 - 1
    g = torch.Generator(device=x.device).manual_seed(7)
    c_txt = torch.randn((1, 77, 128), generator=g, device=x.device).repeat(x.shape[0], 1, 1)
    return dict(c_txt=c_txt, c_img=c_img.contiguous())


def decode_fn(z):
    return torch.sigmoid(F.interpolate(z[:, :3], scale_factor=8, mode="nearest"))


def test_restore_image_is_independent_of_tile_batching(cuda_lib, manifests):
    from tair_b200 import pipeline
    from tair_b200.model.gaussian_diffusion import val_diffusion
    from tair_b200.sampler import SpacedSampler
    m = narrow_model(manifests)
    sampler = SpacedSampler(val_diffusion().betas, "v", False)
    lq = np.random.default_rng(0).integers(0, 256, (200, 300, 3), dtype=np.uint8)   # 2 x 3 tiles
    outs = []
    for tb, graph in ((6, True), (4, True), (1, False)):
        outs.append(pipeline.restore_image(lq, m, sampler, cond_fn=cond_fn, decode_fn=decode_fn, steps=4,
                                           tile_batch=tb, use_cuda_graph=graph))
    assert outs[0].shape == (1, 3, 800, 1200) and torch.isfinite(outs[0]).all()
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])


def test_merge_patches_signature_and_oracle(cuda_lib):
    from oracle import tiles as OT
    from tair_b200.tiles import merge_patches_with_overlap
    g = torch.Generator().manual_seed(3)
    patches = [torch.rand((1, 3, 512, 512), generator=g) for _ in range(6)]
    ref = OT.merge_tiles(patches, (200, 300))
    out = merge_patches_with_overlap([p.cuda() for p in patches], (200, 300), 512, 64)
    assert torch.equal(out.cpu(), ref)


def test_restore_image_full_path_on_kernels(cuda_lib, manifests):
    """LQ pixels -> kernel VAE encoder + kernel CLIP text encoder -> denoise -> kernel VAE decoder -> blend, with the
    default cond / decode stages of restore_image (narrow UNet with a 1024-wide context so that CLIP plugs in);
    batching must not change a bit, and the decoded tiles must match the fp32 oracle VAE on the same latents."""
    import zlib
    from oracle import vae as OV, weights
    from tair_b200 import pipeline
    from tair_b200.init import nondegenerate_init_
    from tair_b200.model import ControlLDM
    from tair_b200.model.gaussian_diffusion import val_diffusion
    from tair_b200.sampler import SpacedSampler
    u = dict(in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[4, 2, 1], num_res_blocks=2,
             channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True,
             use_linear_in_transformer=True, transformer_depth=1, context_dim=1024, legacy=False)
    c = dict(u)
    c.pop("out_channels")
    c["hint_channels"] = 4
    vae_cfg = dict(ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                                 ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0), embed_dim=4)
    clip_cfg = dict(embed_dim=1024, vision_cfg=None, layer="penultimate",
                    text_cfg=dict(context_length=77, vocab_size=49408, width=1024, heads=16, layers=24))
    m = ControlLDM(u, vae_cfg, clip_cfg, c).cuda().eval()
    nondegenerate_init_(m, 5)
    m.vae.load_state_dict(weights.seeded_state_dict(manifests["vae"]))

    def tok(texts):
        out = torch.zeros((len(texts), 77), dtype=torch.long)
        for i, s in enumerate(texts):
            ids = [49406] + [zlib.crc32(w.encode()) % 49000 for w in s.split()][:75] + [49407]
            out[i, :len(ids)] = torch.tensor(ids)
        return out
    m.clip.attach_tokenizer(tok)
    sampler = SpacedSampler(val_diffusion().betas, "v", False)
    lq = np.random.default_rng(1).integers(0, 256, (128, 240, 3), dtype=np.uint8)   # 1 x 2 tiles
    a = pipeline.restore_image(lq, m, sampler, steps=3, tile_batch=2)
    b = pipeline.restore_image(lq, m, sampler, steps=3, tile_batch=1, use_cuda_graph=False)
    assert a.shape == (1, 3, 512, 960) and torch.isfinite(a).all() and 0 <= a.min() and a.max() <= 1
    assert torch.equal(a, b)
    with pytest.raises(RuntimeError):
        pipeline.restore_image(lq, narrow_model(manifests), sampler, steps=1)


def test_gpu_tile_front_end_is_bit_identical_to_pil(cuda_lib):
    """TileFrontEnd (crop + PIL-exact bicubic + /255 on the device) against the reference's host path:
    split_image_with_overlap -> PIL resize(BICUBIC) -> ToTensor, for an image that needs right/bottom padding."""
    from tair_b200 import pipeline, tiles as T
    lq = np.random.default_rng(5).integers(0, 256, (200, 300, 3), dtype=np.uint8)
    pil_tiles = T.split_image_with_overlap(lq, T.LQ_PATCH, T.LQ_OVERLAP)
    ref = torch.stack([pipeline._tile_to_tensor(t) for t in pil_tiles])
    front = T.TileFrontEnd(lq, torch.device("cuda"))
    assert len(front) == len(pil_tiles) == 6
    got = front.tiles(range(len(front)))
    assert got.shape == (6, 3, 512, 512) and torch.equal(got.cpu(), ref)
    sub = front.tiles([4, 1])
    assert torch.equal(sub.cpu(), ref[[4, 1]])

"""Multi-GPU check of the tile-sharded restoration path (BASELINE config 4 geometry: 512x512 LQ -> 25 tiles -> 2048^2).
Run with torchrun.  FULL=1 (default): the whole path on our kernels — VAE encode + CLIP text encoder -> 50-step denoise
(+ TESTR head and per-step prompt re-encoding when TESTR=1) -> VAE decode -> NCCL all-gather -> blend, with the SwinIR
cleaner on the kernels in front (GPU tile front-end -> SwinIR -> VAE encode); the tokenizer is a hash (the CLIP
merge table does not ship here).  FULL=0: cheap
stand-ins for cond / decode (sharding + all-gather + blend only).  Asserts that all ranks (and, with EXPECT_CRC, all
world sizes) produce the same bits."""
import os, sys, time, zlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
import torch.nn.functional as F
from bench import full_cfgs
from tair_b200.init import nondegenerate_init_
from tair_b200.model import ControlLDM
from tair_b200.model.gaussian_diffusion import val_diffusion
from tair_b200.sampler import SpacedSampler
from tair_b200 import pipeline

steps = int(os.environ.get("STEPS", "50"))
world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
FULL = os.environ.get("FULL", "1") == "1"
TESTR = os.environ.get("TESTR", "0") == "1"
unet_cfg, cn_cfg = full_cfgs()
if FULL:
    vae_cfg = dict(ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                                 ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0), embed_dim=4)
    clip_cfg = dict(embed_dim=1024, vision_cfg=None, layer="penultimate",
                    text_cfg=dict(context_length=77, vocab_size=49408, width=1024, heads=16, layers=24))
    model = ControlLDM(unet_cfg, vae_cfg, clip_cfg, cn_cfg).to(dev).eval()
    def hash_tokenizer(texts):
        out = torch.zeros((len(texts), 77), dtype=torch.long)
        for i, s_ in enumerate(texts):
            ids = [49406] + [zlib.crc32(w.encode()) % 49000 for w in s_.split()][:75] + [49407]
            out[i, :len(ids)] = torch.tensor(ids)
        return out
    model.clip.attach_tokenizer(hash_tokenizer)
    from tair_b200.model.swinir import SwinIR
    cleaner = SwinIR(img_size=64, patch_size=1, in_chans=3, embed_dim=180, depths=[6] * 8, num_heads=[6] * 8, window_size=8,
                     mlp_ratio=2, sf=8, img_range=1.0, upsampler="nearest+conv", resi_connection="1conv", unshuffle=True,
                     unshuffle_scale=8).to(dev).eval()
    nondegenerate_init_(cleaner, 77)
else:
    model = ControlLDM(unet_cfg, cn_cfg).to(dev).eval()
nondegenerate_init_(model, 1234)
det = None
if TESTR:
    from types import SimpleNamespace
    from tair_b200.testr import TransformerDetector, default_cfg
    det = TransformerDetector(default_cfg("cuda")).to(dev).eval(); nondegenerate_init_(det, 99)
    model.return_nhwc_feats = True
    vcfg = SimpleNamespace(exp_args=SimpleNamespace(mode="VAL", prompt_style="CAPTION"))
sampler = SpacedSampler(val_diffusion().betas, "v", False)
rng = np.random.default_rng(0)
lq = rng.integers(0, 256, (512, 512, 3), dtype=np.uint8)

def cond_fn(x):   # STAND-IN for SwinIR + VAE-encode + CLIP: deterministic function of the tile pixels
    c_img = F.avg_pool2d(x, 8).mean(1, keepdim=True).repeat(1, 4, 1, 1) * 2 - 1
    g = torch.Generator(device=x.device).manual_seed(7)
    c_txt = torch.randn((1, 77, 1024), generator=g, device=x.device).repeat(x.shape[0], 1, 1)
    return dict(c_txt=c_txt, c_img=c_img.contiguous())

def decode_fn(z):  # STAND-IN for the VAE decoder: nearest x8 of three latent channels squashed to [0,1]
    return torch.sigmoid(F.interpolate(z[:, :3], scale_factor=8, mode="nearest"))

def run(group_world):
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    kw = dict(cleaner=lambda x: cleaner(x).clamp(0, 1)) if FULL else dict(cond_fn=cond_fn, decode_fn=decode_fn)
    if det is not None: kw.update(ts_model=det, cfg=vcfg)
    out = pipeline.restore_image(lq, model, sampler, steps=steps, tile_batch=16, **kw)
    torch.cuda.synchronize()
    return out, time.perf_counter() - t0

out, _ = run(world)          # warm-up (graph capture)
out, dt = run(world)
crc = zlib.crc32(out.cpu().numpy().tobytes())
if world > 1:
    crcs = [None] * world
    dist.all_gather_object(crcs, crc)
    assert len(set(crcs)) == 1, f"ranks disagree on the stitched image: {crcs}"
if rank == 0:
    print(f"full={int(FULL)} testr={int(TESTR)} world={world} tiles=25 steps={steps} out={tuple(out.shape)} crc={crc:08x} time={dt:.3f}s patches/s={25 / dt:.2f}", flush=True)
    exp = os.environ.get("EXPECT_CRC")
    if exp:
        assert f"{crc:08x}" == exp, f"result depends on world size: {crc:08x} vs {exp}"
if world > 1:
    dist.destroy_process_group()

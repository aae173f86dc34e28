"""Worker of tests/test_dist_gpu.py: restore one image with pipeline.restore_image on the narrow ControlLDM + the kernel
VAE, tiles sharded over the ranks of this launch (plain python = 1 rank, torchrun = N ranks), one NCCL
all_gather_into_tensor + the blend kernel; prints ``CRC <hex> world <N>`` from rank 0 after checking that every rank holds
the same image (val_patches.py:316-375 with the tile loop data-parallel over GPUs)."""
import json
import os
import sys
import zlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402
import torch.nn.functional as F  # noqa: E402

from oracle import weights  # noqa: E402  (seeded test weights; tests may import the oracle)
from tair_b200 import pipeline  # noqa: E402
from tair_b200.model import ControlLDM  # noqa: E402
from tair_b200.model.gaussian_diffusion import val_diffusion  # noqa: E402
from tair_b200.sampler import SpacedSampler  # noqa: E402

world = int(os.environ.get("WORLD_SIZE", "1"))
rank = int(os.environ.get("RANK", "0"))
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
man = json.load(open(os.path.join(ROOT, "tests", "golden", "manifests.json")))
u = dict(in_channels=4, out_channels=4, model_channels=64, attention_resolutions=[4, 2, 1], num_res_blocks=2,
         channel_mult=[1, 2, 4, 4], num_head_channels=64, use_spatial_transformer=True, use_linear_in_transformer=True,
         transformer_depth=1, context_dim=128, legacy=False)
c = dict(u)
c.pop("out_channels")
c["hint_channels"] = 4
vae_cfg = dict(ddconfig=dict(double_z=True, z_channels=4, resolution=256, in_channels=3, out_ch=3, ch=128,
                             ch_mult=[1, 2, 4, 4], num_res_blocks=2, attn_resolutions=[], dropout=0.0), embed_dim=4)
m = ControlLDM(u, vae_cfg, None, c)
m.unet.load_state_dict(weights.seeded_state_dict(man["unet_narrow"]))
m.controlnet.load_state_dict(weights.seeded_state_dict(man["controlnet_narrow"]))
m.vae.load_state_dict(weights.seeded_state_dict(man["vae"]))
m = m.to(dev).eval()


def cond_fn(x):   # kernel VAE encoder for c_img; a fixed context stands in for the text encoder (narrow context width)
    g = torch.Generator(device=x.device).manual_seed(7)
    c_txt = torch.randn((1, 77, 128), generator=g, device=x.device).repeat(x.shape[0], 1, 1)
    return dict(c_txt=c_txt, c_img=m.vae_encode(x * 2 - 1, sample=False))


sampler = SpacedSampler(val_diffusion().betas, "v", False)
lq = np.random.default_rng(0).integers(0, 256, (300, 256, 3), dtype=np.uint8)      # 3 x 3 = 9 tiles: ragged over 2 ranks
out = pipeline.restore_image(lq, m, sampler, cond_fn=cond_fn, steps=int(os.environ.get("STEPS", "4")),
                             tile_batch=int(os.environ.get("TILE_BATCH", "4")))
torch.cuda.synchronize()
assert tuple(out.shape) == (1, 3, 1200, 1024) and torch.isfinite(out).all()
crc = zlib.crc32(out.cpu().numpy().tobytes())
if world > 1:
    crcs = [None] * world
    dist.all_gather_object(crcs, crc)
    assert len(set(crcs)) == 1, f"ranks disagree on the stitched image: {crcs}"
if rank == 0:
    print(f"CRC {crc:08x} world {world}", flush=True)
if world > 1:
    dist.destroy_process_group()

"""BASELINE configs[4]: a Real-Text-shaped 4K LQ image (3840x2160 -> 35 x 20 = 700 tiles -> 15360 x 8640 output) with
classifier-free guidance (scale 4.0, cond / uncond stacked as batch 2 per tile), tiles-per-GPU batch sweep, tiles sharded
over the ranks of this launch (torchrun), one all-gather + blend.  TILES=<n> restricts the image to its first n tile rows x
35 columns for a shorter run.  Prints one JSON line per batch size from rank 0.

    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/config4_sweep.py
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist
from bench import CLIP_CFG, SWINIR_CFG, VAE_CFG, full_cfgs, hash_tokenizer
from tair_b200 import pipeline
from tair_b200.init import nondegenerate_init_
from tair_b200.model import ControlLDM
from tair_b200.model.gaussian_diffusion import val_diffusion
from tair_b200.model.swinir import SwinIR
from tair_b200.sampler import SpacedSampler

world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
if world > 1:
    dist.init_process_group("nccl", device_id=dev)
u, c = full_cfgs()
model = ControlLDM(u, VAE_CFG, CLIP_CFG, c).to(dev).eval()
nondegenerate_init_(model, 1234)
model.clip.attach_tokenizer(hash_tokenizer)
cleaner = SwinIR(**SWINIR_CFG).to(dev).eval(); nondegenerate_init_(cleaner, 77)
sampler = SpacedSampler(val_diffusion().betas, "v", False)
rows = int(os.environ.get("TILE_ROWS", "20"))
H = (rows - 1) * 112 + 128 if rows < 20 else 2160
lq = np.random.default_rng(0).integers(0, 256, (H, 3840, 3), dtype=np.uint8)
n_tiles = len(pipeline.T.TileFrontEnd(lq, dev))
uncond_fn = lambda x: model.prepare_condition(cleaner(x).clamp(0, 1), [""] * x.shape[0])
for bsz in [int(v) for v in os.environ.get("BATCHES", "4,16,32").split(",")]:
    def run():
        return pipeline.restore_image(lq, model, sampler, steps=50, tile_batch=bsz, cfg_scale=4.0, uncond_fn=uncond_fn,
                                      cleaner=lambda x: cleaner(x).clamp(0, 1))
    if os.environ.get("WARM", "1") == "1":
        wrows = max(2, -(-world * bsz // 35))          # enough 35-tile rows that every rank sees one full batch of bsz tiles
        pipeline.restore_image(lq[:(wrows - 1) * 112 + 128], model, sampler, steps=2, tile_batch=bsz, cfg_scale=4.0,
                               uncond_fn=uncond_fn, cleaner=lambda x: cleaner(x).clamp(0, 1))   # graph capture / tuning
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    out = run()
    torch.cuda.synchronize()
    if world > 1: dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        print(json.dumps(dict(config="configs[4]", world=world, tiles=n_tiles, tiles_per_gpu_batch=bsz, cfg_scale=4.0, seconds=round(dt, 2),
                              patches_per_s=round(n_tiles / dt, 2), out=list(out.shape))), flush=True)
if world > 1:
    dist.destroy_process_group()

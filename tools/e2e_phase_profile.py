"""Where the time of one pixels-to-pixels image goes (configs[3], pipeline.restore_image, 25 tiles in batches of 16 + 9):
wall clock around every stage with a device synchronize on both sides, plus per-step host/device split of val_sample.
The stages are wrapped from outside (nothing in the product changes); the synchronizes remove the overlap of host and
device work, so the per-stage sum is an upper bound of the un-instrumented time, which is printed beside it."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np
import torch
from bench import BATCH, CLIP_CFG, VAE_CFG, SWINIR_CFG, full_cfgs, hash_tokenizer
from tair_b200 import pipeline
from tair_b200 import tiles as T
from tair_b200.init import nondegenerate_init_
from tair_b200.model import ControlLDM
from tair_b200.model.gaussian_diffusion import val_diffusion
from tair_b200.model.swinir import SwinIR
from tair_b200.sampler import SpacedSampler
from tair_b200.testr import TransformerDetector, default_cfg

dev = torch.device("cuda:0")
u, c = full_cfgs()
model = ControlLDM(u, VAE_CFG, CLIP_CFG, c).to(dev).eval()
nondegenerate_init_(model, 1234)
nondegenerate_init_(model.vae, 1236)
nondegenerate_init_(model.clip, 1237)
model.clip.attach_tokenizer(hash_tokenizer)
det = TransformerDetector(default_cfg("cuda")).to(dev).eval(); nondegenerate_init_(det, 99)
cleaner = SwinIR(**SWINIR_CFG).to(dev).eval(); nondegenerate_init_(cleaner, 77)
sampler = SpacedSampler(val_diffusion().betas, "v", False)
vcfg = SimpleNamespace(exp_args=SimpleNamespace(mode="VAL", prompt_style="CAPTION"))
lq = np.random.default_rng(0).integers(0, 256, (512, 512, 3), dtype=np.uint8)

acc = {}
ON = [False]


def wrap(obj, name, label):
    fn = getattr(obj, name)

    def timed(*a, **k):
        if not ON[0]:
            return fn(*a, **k)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        out = fn(*a, **k)
        torch.cuda.synchronize(); acc.setdefault(label, []).append(1e3 * (time.perf_counter() - t0))
        return out
    setattr(obj, name, timed)


def restore():
    return pipeline.restore_image(lq, model, sampler, steps=50, tile_batch=BATCH, ts_model=det, cfg=vcfg,
                                  cleaner=lambda x: cleaner(x).clamp(0, 1), use_cuda_graph=True).cpu()


restore()                                   # captures
torch.cuda.synchronize()
plain = []
for _ in range(3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    restore()
    torch.cuda.synchronize(); plain.append(round(1e3 * (time.perf_counter() - t0), 1))
m0 = torch.cuda.memory_stats()["num_device_alloc"]
wrap(T.TileFrontEnd, "tiles", "tile front-end (crop + bicubic)")
wrap(cleaner, "forward", "SwinIR")
wrap(model, "prepare_condition", "prepare_condition (VAE encode + CLIP, includes SwinIR)")
wrap(sampler, "val_sample", "val_sample (50 steps)")
wrap(model, "vae_decode", "VAE decode")
wrap(T, "gather_tiles", "gather_tiles")
wrap(T, "merge_patches_with_overlap", "blend")
ON[0] = True
inst = []
for _ in range(2):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    restore()
    torch.cuda.synchronize(); inst.append(round(1e3 * (time.perf_counter() - t0), 1))
out = {"ms_per_image_uninstrumented": plain, "ms_per_image_instrumented": inst,
       "stages_ms (per call, both images)": {k: [round(x, 1) for x in v] for k, v in acc.items()},
       "cudaMallocs_during_instrumented_images": torch.cuda.memory_stats()["num_device_alloc"] - m0,
       "step_graphs": len(sampler._graphs), "clip_graphs": len(model.clip._graphs)}
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open(os.path.join("gpurun_out", "e2e_phase_profile.json"), "w"), indent=1)

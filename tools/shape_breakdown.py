"""Per-shape timing of one ControlNet+UNet denoising step at B=16 (eager, CUDA events around each launch)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import full_cfgs
from tair_b200 import ops
from tair_b200.init import nondegenerate_init_
from tair_b200.model import ControlLDM
from tair_b200.model.gaussian_diffusion import val_diffusion
from tair_b200.sampler import SpacedSampler

B = int(os.environ.get("B", "16"))
dev = torch.device("cuda:0")
model = ControlLDM(*full_cfgs()).to(dev).eval(); nondegenerate_init_(model, 1234)
model.overlap_controlnet = False   # one stream: per-launch event times must be per-kernel times
s = SpacedSampler(val_diffusion().betas, "v", False); s.make_schedule(50); s.to(dev)
g = torch.Generator(device=dev).manual_seed(B)
x = torch.randn((B, 4, 64, 64), device=dev, generator=g)
cond = dict(c_txt=torch.randn((B, 77, 1024), device=dev, generator=g), c_img=torch.randn((B, 4, 64, 64), device=dev, generator=g))
mt = torch.full((B,), 500, device=dev, dtype=torch.long); tt = torch.full((B,), 25, device=dev, dtype=torch.long)
nz = torch.randn_like(x)
for _ in range(5): s.p_sample(model, x, mt, tt, cond, None, 1.0, noise=nz)
t = ops.KernelTimer(); ops.set_timer(t)
R = 5
for _ in range(R): s.p_sample(model, x, mt, tt, cond, None, 1.0, noise=nz)
ops.set_timer(None)
rows = sorted(t.by_shape().items(), key=lambda kv: -kv[1]["ms"])
tot = sum(v["ms"] for _, v in rows) / R
print(f"total timed {tot:.2f} ms/step")
for (fam, tag), v in rows[:60]:
    ms = v["ms"] / R
    rate = v["work"] / v["ms"] / 1e9
    unit = "GB/s" if fam in ("groupnorm", "layernorm") else "TF/s"
    print(f"{fam:10s} {str(tag):42s} x{v['launches'] // R:3d} {ms:7.3f} ms {100 * ms / tot:5.1f}%  {rate:8.0f} {unit} ({ms / (v['launches'] // R) * 1e3:.1f} us each)")

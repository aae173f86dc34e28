"""One eager denoising step (B=16, full config) bracketed by cudaProfilerStart/Stop for ncu
(`--profile-from-start off`).  Prints nothing performance-related: numbers under ncu are never bench values."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import full_cfgs, BATCH
from tair_b200.init import nondegenerate_init_
from tair_b200.model import ControlLDM
from tair_b200.model.gaussian_diffusion import val_diffusion
from tair_b200.sampler import SpacedSampler

B = int(sys.argv[1]) if len(sys.argv) > 1 else BATCH
dev = torch.device("cuda:0")
model = ControlLDM(*full_cfgs()).to(dev).eval()
nondegenerate_init_(model, 1234)
model.overlap_controlnet = False   # one stream: ncu serialises kernels anyway, keep the launch order canonical
model.return_nhwc_feats = True     # as SpacedSampler.sample / val_sample run it: decoder features stay channels-last bf16
s = SpacedSampler(val_diffusion().betas, "v", False)
s.make_schedule(50); s.to(dev)
g = torch.Generator(device=dev).manual_seed(100)
x = torch.randn((B, 4, 64, 64), device=dev, generator=g)
cond = dict(c_txt=torch.randn((B, 77, 1024), device=dev, generator=g), c_img=torch.randn((B, 4, 64, 64), device=dev, generator=g))
mt = torch.full((B,), 500, device=dev, dtype=torch.long); tt = torch.full((B,), 25, device=dev, dtype=torch.long)
for _ in range(2):
    s.p_sample(model, x, mt, tt, cond, None, 1.0)
torch.cuda.synchronize()
torch.cuda.profiler.start()
out, _ = s.p_sample(model, x, mt, tt, cond, None, 1.0)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
assert torch.isfinite(out).all()
print("step ok")

"""TESTR head (dense part: projections, deformable encoder, proposals, both decoders, prediction heads) on one step's
decoder features at B=16: CUDA-graph replay time of the whole head, and the eager per-shape table of OUR kernels
(CUDA events around each launch; whatever the sum leaves of the replay time is torch glue: top-k, gathers, sigmoids)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import BATCH, full_cfgs
from tair_b200 import ops
from tair_b200.init import nondegenerate_init_
from tair_b200.model import ControlLDM
from tair_b200.model.gaussian_diffusion import val_diffusion
from tair_b200.sampler import SpacedSampler
from tair_b200.testr import TransformerDetector, default_cfg

B = int(os.environ.get("B", str(BATCH)))
dev = torch.device("cuda:0")
model = ControlLDM(*full_cfgs()).to(dev).eval(); nondegenerate_init_(model, 1234)
det = TransformerDetector(default_cfg("cuda")).to(dev).eval(); nondegenerate_init_(det, 99)
s = SpacedSampler(val_diffusion().betas, "v", False); s.make_schedule(50); s.to(dev)
g = torch.Generator(device=dev).manual_seed(B)
x = torch.randn((B, 4, 64, 64), device=dev, generator=g)
cond = dict(c_txt=torch.randn((B, 77, 1024), device=dev, generator=g), c_img=torch.randn((B, 4, 64, 64), device=dev, generator=g))
mt = torch.full((B,), 500, device=dev, dtype=torch.long); tt = torch.full((B,), 25, device=dev, dtype=torch.long)
model.return_nhwc_feats = True
_, feats = s.p_sample(model, x, mt, tt, cond, None, 1.0, noise=torch.randn_like(x))
for _ in range(3): det.testr(feats)
torch.cuda.synchronize()
side = torch.cuda.Stream()
side.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(side):
    det.testr(feats)
torch.cuda.current_stream().wait_stream(side)
gr = torch.cuda.CUDAGraph()
with torch.cuda.graph(gr, stream=side):
    out = det.testr(feats)
for _ in range(3): gr.replay()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
R = 10
a.record()
for _ in range(R): gr.replay()
b.record(); torch.cuda.synchronize()
print(f"TESTR dense head, B={B}: {a.elapsed_time(b) / R:.3f} ms per graph replay")
ops.reset_launch_count(); det.testr(feats); torch.cuda.synchronize()
print("launches of our kernels per head call:", ops.launch_count())
t = ops.KernelTimer(); ops.set_timer(t)
R = 5
for _ in range(R): det.testr(feats)
ops.set_timer(None)
rows = sorted(t.by_shape().items(), key=lambda kv: -kv[1]["ms"])
tot = sum(v["ms"] for _, v in rows) / R
print(f"sum over our kernels (eager, ~5 us high per short launch): {tot:.2f} ms")
fam = {}
for (f, tag), v in rows:
    fam[f] = fam.get(f, 0.0) + v["ms"] / R
print({k: round(v, 3) for k, v in sorted(fam.items(), key=lambda kv: -kv[1])})
for (f, tag), v in rows[:45]:
    ms = v["ms"] / R
    print(f"{f:12s} {str(tag):48s} x{v['launches'] // R:3d} {ms:7.3f} ms {100 * ms / tot:5.1f}%  ({ms / (v['launches'] // R) * 1e3:.1f} us each)")

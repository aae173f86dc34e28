"""Where the time of one FULL TeReDiff step goes at B=16 (configs[2]): graph replay (ControlNet + UNet + update + dense
TESTR head) vs detection post-processing vs string decode vs prompt vs tokenizer vs CLIP vs conditioning copy.
Wall-clock with a synchronize after every phase (so phases do not overlap): the per-phase sum is an upper bound."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from types import SimpleNamespace
import numpy as np
import torch
from bench import BATCH, CLIP_CFG, VAE_CFG, full_cfgs, hash_tokenizer
from tair_b200.init import nondegenerate_init_
from tair_b200.model import ControlLDM
from tair_b200.model.gaussian_diffusion import val_diffusion
from tair_b200.prompt import build_prompt, decode_texts
from tair_b200.sampler import SpacedSampler
from tair_b200.testr import TransformerDetector, default_cfg

dev = torch.device("cuda:0")
u, c = full_cfgs()
model = ControlLDM(u, None, CLIP_CFG, c).to(dev).eval()
nondegenerate_init_(model, 1234)
model.clip.attach_tokenizer(hash_tokenizer)
det = TransformerDetector(default_cfg("cuda")).to(dev).eval(); nondegenerate_init_(det, 99)
s = SpacedSampler(val_diffusion().betas, "v", False)
B = BATCH
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((B, 4, 64, 64), device=dev, generator=g)
cond = dict(c_txt=model.clip.encode([""] * B), c_img=torch.randn((B, 4, 64, 64), device=dev, generator=g))
s.make_schedule(50); s.to(dev)
model.return_nhwc_feats = True
st = s._stepper("val", model, x, cond, None, 1.0, head=det.testr, extra=(True,))
nz = torch.randn_like(x)
acc = {}
def phase(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize(); acc[name] = acc.get(name, 0.0) + (time.perf_counter() - t0)
    return out
N = 12
for it in range(N + 2):
    if it == 2: acc.clear()
    st.load_cond(cond, None)
    xo, feats, dense = phase("graph (CN+UNet+update+TESTR dense)", lambda: st.run(x, 500, 25, nz, 1.0))
    res = phase("inference (threshold/gather)", lambda: det.inference(dense["pred_logits"], dense["pred_ctrl_points"], dense["pred_texts"], [(512, 512)] * B))
    texts, polys = phase("decode_texts (D2H + strings)", lambda: decode_texts(res))
    prompts = phase("build_prompt", lambda: [build_prompt(t, "CAPTION") for t in texts])
    # make every step's prompts new, as in a real run where detections change
    prompts = [p + f" {it}" for p in prompts]
    toks = phase("tokenizer", lambda: hash_tokenizer(prompts))
    cond["c_txt"] = phase("clip.encode (new prompts)", lambda: model.clip.encode(prompts))
tot = sum(acc.values())
out = {k: round(1e3 * v / N, 3) for k, v in acc.items()}
out["sum_ms"] = round(1e3 * tot / N, 3)
out["detections_tile0"] = len(texts[0])
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/full_step_profile.json", "w"), indent=1)

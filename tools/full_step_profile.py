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
OLD = "--old" in sys.argv
g = torch.Generator(device=dev).manual_seed(0)
x = torch.randn((B, 4, 64, 64), device=dev, generator=g)
cond = dict(c_txt=model.clip.encode([""] * B), c_img=torch.randn((B, 4, 64, 64), device=dev, generator=g))
s.make_schedule(50); s.to(dev)
model.return_nhwc_feats = True
st = s._stepper("val", model, x, cond, None, 1.0, head=det.testr, extra=(True,))
nz = torch.randn_like(x)
acc, per_it = {}, {}
import gc
gc_log, gc_t0 = [], [0.0]
def _gc_cb(ph, info):   # garbage-collector pauses, by generation: they land in whichever phase allocates
    if ph == "start": gc_t0[0] = time.perf_counter()
    else: gc_log.append((info["generation"], round(1e3 * (time.perf_counter() - gc_t0[0]), 2)))
gc.callbacks.append(_gc_cb)
if "--no-gc" in sys.argv: gc.disable()
def phase(name, fn):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = fn()
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    acc[name] = acc.get(name, 0.0) + dt
    per_it.setdefault(name, []).append(round(1e3 * dt, 2))
    return out
N = 12
for it in range(N + 2):
    if it == 2: acc.clear(); per_it.clear(); gc_log.clear(); mallocs0 = torch.cuda.memory_stats()['num_device_alloc']
    st.load_cond(cond, None)
    xo, feats, dense = phase("graph (CN+UNet+update+TESTR dense)", lambda: st.run(x, 500, 25, nz, 1.0))
    if OLD:   # the round-1 path: eager softmax / threshold / gathers, then one D2H per tile and Python string building
        res = phase("inference (threshold/gather)", lambda: det.inference(dense["pred_logits"], dense["pred_ctrl_points"], dense["pred_texts"], [(512, 512)] * B))
        texts, polys = phase("decode_texts (D2H + strings)", lambda: decode_texts(res))
    else:     # what val_sample calls: tair_testr_postprocess + one D2H + numpy decode
        texts, polys = phase("detect_host (1 kernel + 1 D2H + numpy decode)", lambda: det.detect_host(dense, (512, 512)))
    prompts = phase("build_prompt", lambda: [build_prompt(t, "CAPTION") for t in texts])
    # make every step's prompts new, as in a real run where detections change
    prompts = [p + f" {it}" for p in prompts]
    toks = phase("tokenizer", lambda: hash_tokenizer(prompts))
    cond["c_txt"] = phase("clip.encode (new prompts)", lambda: model.clip.encode(prompts))
tot = sum(acc.values())
out = {k: round(1e3 * v / N, 3) for k, v in acc.items()}
out["sum_ms"] = round(1e3 * tot / N, 3)
out["detections_tile0"] = len(texts[0])
out["distinct_prompts_last_step"] = len(set(prompts))
out["clip_graph_keys"] = [k[:2] for k in model.clip._graphs]
out["gc_pauses_ms_by_generation"] = {g: [t for gg, t in gc_log if gg == g and t > 0.2] for g in (0, 1, 2)}
out["cudaMallocs_in_timed_iterations"] = torch.cuda.memory_stats()["num_device_alloc"] - mallocs0
out["per_iteration_ms"] = per_it
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/full_step_profile.json", "w"), indent=1)

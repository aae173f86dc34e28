"""Per-step latency of the full hot path (BASELINE configs 2/3): ControlNet+UNet+update, and with the TESTR head,
eager vs CUDA-graph, B in {1, 16}.  CUDA-event timing, 3 warm-up + 10 timed replays."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import full_cfgs
from tair_b200 import ops
from tair_b200.init import nondegenerate_init_
from tair_b200.model import ControlLDM
from tair_b200.model.gaussian_diffusion import val_diffusion
from tair_b200.sampler import SpacedSampler
from tair_b200.sampler.spaced_sampler import _StepGraph
from tair_b200.testr import TransformerDetector, default_cfg

dev = torch.device("cuda:0")
model = ControlLDM(*full_cfgs()).to(dev).eval(); nondegenerate_init_(model, 1234)
det = TransformerDetector(default_cfg("cuda")).to(dev).eval(); nondegenerate_init_(det, 99)
s = SpacedSampler(val_diffusion().betas, "v", False); s.make_schedule(50); s.to(dev)

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n

from tair_b200.model.clip import FrozenOpenCLIPEmbedder
clip = FrozenOpenCLIPEmbedder(1024, None, dict(context_length=77, vocab_size=49408, width=1024, heads=16, layers=24),
                              layer="penultimate").to(dev).eval(); nondegenerate_init_(clip, 7)
import zlib
def hash_tok(texts):
    out = torch.zeros((len(texts), 77), dtype=torch.long)
    for i, s_ in enumerate(texts):
        ids = [49406] + [zlib.crc32(w.encode()) % 49000 for w in s_.split()][:75] + [49407]
        out[i, :len(ids)] = torch.tensor(ids)
    return out
clip.attach_tokenizer(hash_tok)

res = {}
for B in (1, 16):
    g = torch.Generator(device=dev).manual_seed(B)
    x = torch.randn((B, 4, 64, 64), device=dev, generator=g)
    cond = dict(c_txt=torch.randn((B, 77, 1024), device=dev, generator=g), c_img=torch.randn((B, 4, 64, 64), device=dev, generator=g))
    mt = torch.full((B,), 500, device=dev, dtype=torch.long); tt = torch.full((B,), 25, device=dev, dtype=torch.long)
    nz = torch.randn_like(x)
    model.return_nhwc_feats = True
    def eager_unet(): return s.p_sample(model, x, mt, tt, cond, None, 1.0, noise=nz)
    def eager_full():
        xp, f = s.p_sample(model, x, mt, tt, cond, None, 1.0, noise=nz)
        return det.testr(f)
    ops.reset_launch_count(); eager_full(); torch.cuda.synchronize(); launches = ops.launch_count()
    r = dict(launches_full_step=launches, eager_unet_ms=timeit(eager_unet), eager_full_ms=timeit(eager_full))
    g1 = _StepGraph(s, model, x, cond, None); g1.run(x, 500, 25, nz, 1.0)
    r["graph_unet_ms"] = timeit(lambda: g1.run(x, 500, 25, nz, 1.0))
    g2 = _StepGraph(s, model, x, cond, None, head=det.testr); g2.run(x, 500, 25, nz, 1.0)
    r["graph_full_ms"] = timeit(lambda: g2.run(x, 500, 25, nz, 1.0))
    r["testr_ms_in_graph"] = r["graph_full_ms"] - r["graph_unet_ms"]
    # inference post-processing (dynamic shapes + D2H): eager
    dense = det.testr(eager_unet()[1])
    r["inference_ms"] = timeit(lambda: [len(q) for q in det.inference(dense["pred_logits"], dense["pred_ctrl_points"], dense["pred_texts"], [(512, 512)] * B)])
    # prompt re-encoding: B distinct new prompts through the 23 causal CLIP blocks (cache cleared = worst case)
    cnt = [0]
    def clip_new():
        cnt[0] += 1
        clip.clear_cache()
        return clip.encode([f"A realistic scene where the texts word{cnt[0]} tile{i} appear clearly" for i in range(B)])
    r["clip_encode_new_prompts_ms"] = timeit(clip_new)
    r["full_step_ms_incl_inference_and_clip"] = r["graph_full_ms"] + r["inference_ms"] + r["clip_encode_new_prompts_ms"]
    res[f"B{B}"] = {k: (round(v, 3) if isinstance(v, float) else v) for k, v in r.items()}
    print(B, res[f"B{B}"], flush=True)
json.dump(res, open("gpurun_out/step_times.json", "w"), indent=1)

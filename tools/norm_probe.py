"""Norm kernels in isolation, CUDA-graph replay of 20 back-to-back launches on rotating buffers (3 x tensor > L2 for the
large shapes), microseconds per call and GB/s on ALGORITHMIC bytes (4 B per element).  Env knobs are read once per process:
TAIR_GN_SLABS, TAIR_GN_UNROLL, TAIR_RS_WARPS, TAIR_LN_WARPS."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)

def bench(fn_list, reps=5):
    for f in fn_list: f()
    torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for f in fn_list: f()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(reps): gr.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / len(fn_list) * 1e3

out = {}
for shp in ((16, 64, 64, 320), (16, 32, 32, 640), (16, 16, 16, 1280), (16, 8, 8, 1280), (16, 64, 64, 640), (16, 16, 16, 2560), (16, 64, 64, 960)):
    xs = [torch.randn(shp, device=dev, generator=g).bfloat16() for _ in range(4)]
    ys = [torch.empty_like(x) for x in xs]
    C = shp[-1]
    ga, be = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    us = bench([(lambda i=i: ops.groupnorm(xs[i % 4], ga, be, act=ops.ACT_SILU, out=ys[i % 4])) for i in range(20)])
    n = xs[0].numel()
    out[f"gn{shp}"] = (round(us, 2), round(4 * n / us / 1e3, 0))
for (M, C) in ((65536, 320), (16384, 640), (4096, 1280)):
    xs = [torch.randn((M, C), device=dev, generator=g).bfloat16() for _ in range(4)]
    ys = [torch.empty_like(x) for x in xs]
    ga, be = torch.ones(C, device=dev), torch.zeros(C, device=dev)
    us_ln = bench([(lambda i=i: ops.layernorm(xs[i % 4], ga, be, out=ys[i % 4])) for i in range(20)])
    us_rs = bench([(lambda i=i: ops.row_stats(xs[i % 4])) for i in range(20)])
    out[f"ln({M},{C})"] = dict(layernorm_us=round(us_ln, 2), row_stats_us=round(us_rs, 2), row_stats_gbs_algorithmic=round(4 * M * C / us_rs / 1e3))
print(json.dumps({k: v for k, v in os.environ.items() if k.startswith("TAIR_")}), json.dumps(out))

"""TESTR's K=256 GEMMs with a per-row positional add (tair_epilogue.rowgroup) in isolation: how much of their time is the
fp32 row add?  CUDA-graph replay of 10 launches."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
def timeit(fn, n=10, reps=5):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n): fn()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(reps): gr.replay()
    b.record(); torch.cuda.synchronize()
    return round(a.elapsed_time(b) / reps / n * 1e3, 1)
res = {}
for (M, N, K, rpg, G) in ((151552, 384, 256, -9472, 9472), (40000, 1536, 256, 25, 1600), (25600, 1536, 256, 16, 1600), (151552, 256, 256, 0, 0), (151552, 1024, 256, 0, 0)):
    a = torch.randn(M, K, device=dev, generator=g).bfloat16()
    w = (torch.randn(N, K, device=dev, generator=g) / 16).bfloat16()
    bias = torch.randn(N, device=dev, generator=g)
    out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
    r = dict(plain=timeit(lambda: ops.gemm(a, w, out=out)), bias=timeit(lambda: ops.gemm(a, w, bias=bias, out=out)))
    if rpg:
        rg = torch.randn(G, N, device=dev, generator=g)
        r["rowgroup_fp32"] = timeit(lambda: ops.gemm(a, w, rowgroup=rg, rows_per_group=rpg, out=out))
    r["hbm_floor_us"] = round((M * K * 2 + M * N * 2) / 6.5517e6, 1)
    res[f"{M}x{N}x{K}"] = r
print(json.dumps(res))

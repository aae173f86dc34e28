"""Key/value-stationary attention in isolation: time vs query tiles per CTA (TAIR_KVS_CHUNKS) -> fixed cost per CTA and
cost per tile.  CUDA-graph replay of 10 launches on rotating buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
res = {}
for (B, H, Lq, Lk) in ((16, 5, 4096, 77), (16, 10, 1024, 77), (16, 20, 256, 77)):
    C = H * 64
    qs = [torch.randn(B * Lq, C, device=dev, generator=g).bfloat16() for _ in range(4)]
    k = torch.randn(B * Lk, C, device=dev, generator=g).bfloat16(); v = torch.randn(B * Lk, C, device=dev, generator=g).bfloat16()
    outs = [torch.empty_like(q) for q in qs]
    def run():
        for i in range(12): ops.attention(qs[i % 4], k, v, B=B, H=H, Lq=Lq, Lk=Lk, out=outs[i % 4])
    run(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr): run()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(5): gr.replay()
    b.record(); torch.cuda.synchronize()
    res[(B, H, Lq, Lk)] = round(a.elapsed_time(b) / 5 / 12 * 1e3, 2)
print("KVS=" + os.environ.get("TAIR_ATTN_KVS", "1"), "CHUNKS=" + os.environ.get("TAIR_KVS_CHUNKS", "auto"), res)

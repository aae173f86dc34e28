"""Self-attention kernel in isolation against torch SDPA (the library it replaces) on the same box: CUDA-graph replay of 10
launches, microseconds per launch and TFLOP/s (4 B H Lq Lk 64 flop)."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
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
    return a.elapsed_time(b) / reps / n * 1e3
res = {}
for (B, H, L) in ((16, 5, 4096), (16, 10, 1024), (16, 20, 256)):
    C = H * 64
    qkv = torch.randn(B * L, 3 * C, device=dev, generator=g).bfloat16()
    out = torch.empty(B * L, C, device=dev, dtype=torch.bfloat16)
    us = timeit(lambda: ops.attention(qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:], B=B, H=H, Lq=L, Lk=L, out=out))
    q, k, v = (qkv[:, i * C:(i + 1) * C].reshape(B, L, H, 64).transpose(1, 2).contiguous() for i in range(3))
    us_t = timeit(lambda: F.scaled_dot_product_attention(q, k, v))
    fl = 4.0 * B * H * L * L * 64
    res[f"B{B}_H{H}_L{L}"] = dict(tair_us=round(us, 1), tair_tflops=round(fl / us / 1e6), torch_sdpa_us=round(us_t, 1), torch_tflops=round(fl / us_t / 1e6))
print(json.dumps(res))

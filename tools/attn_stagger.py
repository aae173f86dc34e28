import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
def bench(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for (B, H, L) in [(16, 5, 4096), (16, 10, 1024), (16, 20, 256)]:
    C = H * 64
    q = torch.randn(B * L, C, device="cuda").bfloat16(); k = torch.randn(B * L, C, device="cuda").bfloat16(); v = torch.randn(B * L, C, device="cuda").bfloat16()
    out = torch.empty(B * L, C, device="cuda", dtype=torch.bfloat16)
    us = bench(lambda: ops.attention(q, k, v, B=B, H=H, Lq=L, Lk=L, out=out))
    print(f"stagger={os.environ.get('TAIR_ATTN_STAGGER', '0')} B{B} H{H} L{L}: {us:.1f} us {4.0 * B * H * L * L * 64 / us / 1e6:.0f} TF/s", flush=True)

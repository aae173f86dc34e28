import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
def bench(fn, n=40):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for (B, HW, C) in [(16, 4096, 320), (16, 1024, 640), (16, 256, 1280), (16, 4096, 640)]:
    xs = [torch.randn(B, HW, C, device="cuda").bfloat16() for _ in range(3)]; g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    line = f"GN+SiLU {B}x{HW}x{C}:"
    for slabs in (64, 32, 16, 8):
        os.environ["TAIR_GN_SLABS"] = str(slabs)
        i = [0]
        def f():
            i[0] = (i[0] + 1) % 3
            ops.groupnorm(xs[i[0]], g, b, groups=32, eps=1e-5, act=ops.ACT_SILU)
        line += f" slabs{slabs}={bench(f):.1f}us"
    print(line, flush=True)

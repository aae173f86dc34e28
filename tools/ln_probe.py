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
for (M, C) in [(65536, 320), (16384, 640), (4096, 1280), (151552, 256)]:
    xs = [torch.randn(M, C, device="cuda").bfloat16() for _ in range(3)]; g = torch.ones(C, device="cuda"); b = torch.zeros(C, device="cuda")
    outs = [torch.empty_like(x) for x in xs]
    line = f"LN {M}x{C}:"
    for wps in (2, 4, 8, 16):
        os.environ["TAIR_LN_WARPS"] = str(wps)
        i = [0]
        def f():
            i[0] = (i[0] + 1) % 3
            ops.layernorm(xs[i[0]], g, b, out=outs[i[0]])
        line += f" warps{wps}={bench(f):.1f}us"
    print(line, flush=True)

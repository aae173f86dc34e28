"""Same-shape comparison of the implicit-GEMM conv and a plain GEMM (M=B*H*W, N=Cout, K=9*Cin) under the kernel's
bring-up probes (TAIR_GEMM_DEBUG: 2 no epilogue, 4/8 load B/A once) to separate MMA, load and epilogue cost."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops

def bench(fn, n=30):
    for _ in range(10): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

# spin the clocks up
z = torch.randn(8192, 8192, device="cuda").bfloat16()
for _ in range(20): z @ z
torch.cuda.synchronize()
for (B, H, Cin, Cout) in [(16, 64, 320, 320), (16, 32, 640, 640), (16, 16, 1280, 1280)]:
    x = torch.randn(B, H, H, Cin, device="cuda").bfloat16(); w = torch.randn(Cout, 9 * Cin, device="cuda").bfloat16()
    out = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.bfloat16)
    a2 = torch.randn(B * H * H, 9 * Cin, device="cuda").bfloat16()
    gf = 2.0 * B * H * H * Cout * 9 * Cin / 1e9
    for name, fn in (("conv", lambda: ops.conv3x3(x, w, out=out)), ("gemm", lambda: ops.gemm(a2, w, out=out.view(-1, Cout)))):
        res = []
        for dbg in (0, 2, 4, 8, 12, 14):
            os.environ["TAIR_GEMM_DEBUG"] = str(dbg)
            us = bench(fn)
            res.append(f"dbg{dbg}={us:.1f}us({gf / us * 1e3 / 1e3:.0f}TF)")
        print(f"{name} {H}x{H} {Cin}->{Cout}", " ".join(res), flush=True)

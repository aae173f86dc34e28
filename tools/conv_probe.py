import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
def bench(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
two = os.environ.get("TAIR_GEMM_2CTA", "1")
for (B, H, Cin, Cout) in [(16, 64, 320, 320), (16, 32, 640, 640), (16, 16, 1280, 1280)]:
    x = torch.randn(B, H, H, Cin, device="cuda").bfloat16(); w = torch.randn(Cout, 9 * Cin, device="cuda").bfloat16()
    out = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.bfloat16)
    res = []
    for dbg in (0, 2, 4, 8, 12, 14):
        os.environ["TAIR_GEMM_DEBUG"] = str(dbg)
        res.append(f"dbg{dbg}={bench(lambda: ops.conv3x3(x, w, out=out)):.1f}us")
    print(f"2cta={two} conv {H}x{H} {Cin}->{Cout}", " ".join(res), flush=True)
M, N, K = 8192, 8192, 8192
a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16(); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
res = []
for dbg in (0, 2, 4, 8, 12, 14):
    os.environ["TAIR_GEMM_DEBUG"] = str(dbg)
    res.append(f"dbg{dbg}={bench(lambda: ops.gemm(a, w, out=out), 5):.1f}us")
print(f"2cta={two} gemm 8192^3", " ".join(res), flush=True)

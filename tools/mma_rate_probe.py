import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
def bench(fn, n=5):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n
os.environ["TAIR_GEMM_2CTA"] = "0"
M, K = 8192 * 2, 4096
for bn in (64, 96, 128, 160, 192, 224, 256):
    N = bn * 32
    a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16(); out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    os.environ["TAIR_GEMM_BN"] = str(bn)
    r = []
    for dbg in (14, 46, 0):
        os.environ["TAIR_GEMM_DEBUG"] = str(dbg)
        ms = bench(lambda: ops.gemm(a, w, out=out))
        r.append(f"dbg{dbg}: {ms*1e3:.0f}us {2.0*M*N*K/ms/1e9:.0f} TF/s")
    print(f"BN={bn} N={N}", " | ".join(r), flush=True)

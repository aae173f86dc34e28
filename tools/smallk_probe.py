import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
os.environ["TAIR_AUTOTUNE"] = "0"
def bench(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3
for (M, N, K) in [(151552, 256, 256), (65536, 320, 320), (151552, 1024, 256)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = (torch.randn(N, K, device="cuda") * 0.05).bfloat16()
    outs = [torch.empty(M, N, device="cuda", dtype=torch.bfloat16) for _ in range(3)]
    for bn in (128, 160, 256):
        os.environ["TAIR_GEMM_BN"] = str(bn)
        line = f"{M}x{N}x{K} BN{bn}:"
        for dbg in (0, 1, 2, 4, 8, 12, 14):
            os.environ["TAIR_GEMM_DEBUG"] = str(dbg)
            i = [0]
            def f():
                i[0] = (i[0] + 1) % 3
                ops.gemm(a, w, out=outs[i[0]])
            line += f" dbg{dbg}={bench(f):.1f}"
        print(line, flush=True)

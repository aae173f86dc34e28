"""Bring-up probe: where does the time of a small-K GEMM go?  (TAIR_GEMM_DEBUG bit flags, see gemm_tc.cu)"""
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

for (M, N, K) in [(65536, 320, 320), (65536, 960, 320), (65536, 2560, 320), (16384, 640, 640), (16384, 1920, 640), (4096, 1280, 1280), (65536, 320, 1280)]:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
    out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
    res = []
    for dbg in (0, 1, 2, 4, 6):
        os.environ["TAIR_GEMM_DEBUG"] = str(dbg)
        res.append(f"dbg{dbg}={bench(lambda: ops.gemm(a, w, out=out)):.1f}us")
    os.environ["TAIR_GEMM_DEBUG"] = "0"
    g = torch.cuda.CUDAGraph()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        ops.gemm(a, w, out=out)
        with torch.cuda.graph(g):
            for _ in range(20): ops.gemm(a, w, out=out)
    torch.cuda.synchronize()
    t = bench(lambda: g.replay(), 5) / 20
    res.append(f"graph={t:.1f}us")
    res.append(f"cublas={bench(lambda: torch.matmul(a, w.t(), out=out)):.1f}us")
    print(M, N, K, " ".join(res), flush=True)

"""Sweep the N-tile width (TAIR_GEMM_BN) over the conv / GEMM shapes of one B=16 denoising step; prints us per launch.
Each measurement: 5 warm-up + 20 timed launches back to back (CUDA events), rotating over 4 output buffers."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops

BNS = (64, 96, 128, 160, 192, 224, 256)
CONVS = [(16, 64, 320, 320, 1), (16, 16, 1280, 1280, 1), (16, 8, 1280, 1280, 1), (16, 32, 640, 640, 1), (16, 16, 2560, 1280, 1),
         (16, 64, 640, 320, 1), (16, 64, 640, 640, 1), (16, 8, 2560, 1280, 1), (16, 32, 1280, 1280, 1), (16, 64, 960, 320, 1),
         (16, 32, 1920, 640, 1), (16, 32, 1280, 640, 1), (16, 16, 1920, 1280, 1), (16, 32, 960, 640, 1), (16, 16, 640, 1280, 1),
         (16, 16, 1280, 1280, 2), (16, 32, 320, 640, 1), (16, 32, 640, 640, 2), (16, 64, 320, 320, 2), (16, 64, 64, 320, 1)]
GEMMS = [(65536, 320, 320), (16384, 640, 640), (4096, 1280, 1280), (65536, 960, 320), (65536, 320, 1280), (16384, 640, 2560),
         (4096, 1280, 5120), (16384, 1920, 640), (4096, 3840, 1280), (1024, 1280, 1280), (65536, 320, 640), (4096, 1280, 2560),
         (1024, 1280, 5120), (1024, 1280, 2560), (16384, 640, 320), (1232, 24960, 1024), (1024, 3840, 1280), (16384, 640, 1280)]

def bench(fn, n=20):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / n * 1e3

def sweep(name, fn):
    os.environ.pop("TAIR_GEMM_BN", None)
    base = bench(fn)
    res = []
    for bn in BNS:
        os.environ["TAIR_GEMM_BN"] = str(bn)
        res.append(bench(fn))
    os.environ.pop("TAIR_GEMM_BN", None)
    best = min(range(len(BNS)), key=lambda i: res[i])
    print(f"{name:34s} auto={base:6.1f} best=BN{BNS[best]}:{res[best]:6.1f} | " + " ".join(f"{bn}:{r:.1f}" for bn, r in zip(BNS, res)), flush=True)

z = torch.randn(8192, 8192, device="cuda").bfloat16()
for _ in range(10): z @ z
for (B, H, Cin, Cout, st) in CONVS:
    x = torch.randn(B, H, H, Cin, device="cuda").bfloat16(); w = torch.randn(Cout, 9 * Cin, device="cuda").bfloat16() * 0.02
    sweep(f"conv {H}x{H} {Cin}->{Cout} s{st}", lambda: ops.conv3x3(x, w, stride=st))
for (M, N, K) in GEMMS:
    a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16() * 0.02
    sweep(f"gemm {M}x{N}x{K}", lambda: ops.gemm(a, w))

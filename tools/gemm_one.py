"""One GEMM shape for Nsight Compute: python tools/gemm_one.py M N K [bias] [res] [relu]; the profiled range (use
--profile-from-start off) holds exactly one launch of the tuned tile."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4])
flags = sys.argv[4:]
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
a = torch.randn(M, K, device=dev, generator=g).bfloat16()
w = (torch.randn(N, K, device=dev, generator=g) / 16).bfloat16()
bias = torch.randn(N, device=dev, generator=g) if "bias" in flags else None
res = torch.randn(M, N, device=dev, generator=g).bfloat16() if "res" in flags else None
out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
act = ops.ACT_RELU if "relu" in flags else ops.ACT_NONE
for _ in range(4):
    ops.gemm(a, w, bias=bias, residual=res, act=act, out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
e0.record()
for _ in range(10):
    ops.gemm(a, w, bias=bias, residual=res, act=act, out=out)
e1.record(); torch.cuda.synchronize()
print(f"{M}x{N}x{K} {flags}: {e0.elapsed_time(e1) * 100:.1f} us per launch (eager back to back)")
torch.cuda.profiler.start()
ops.gemm(a, w, bias=bias, residual=res, act=act, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()

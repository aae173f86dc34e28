import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
M, N, K = (int(v) for v in sys.argv[1:4]) if len(sys.argv) > 3 else (65536, 320, 320)
a = torch.randn(M, K, device="cuda").bfloat16(); w = torch.randn(N, K, device="cuda").bfloat16()
out = torch.empty(M, N, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.gemm(a, w, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.gemm(a, w, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")

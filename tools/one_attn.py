import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
B, H, L = 16, 5, 4096
C = H * 64
q = torch.randn(B * L, C, device="cuda").bfloat16(); k = torch.randn(B * L, C, device="cuda").bfloat16(); v = torch.randn(B * L, C, device="cuda").bfloat16()
out = torch.empty(B * L, C, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.attention(q, k, v, B=B, H=H, Lq=L, Lk=L, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.attention(q, k, v, B=B, H=H, Lq=L, Lk=L, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")

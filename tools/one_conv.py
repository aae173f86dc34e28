import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
B, H, Cin, Cout = (int(v) for v in (sys.argv[1:5] if len(sys.argv) > 4 else (16, 64, 320, 320)))
x = torch.randn(B, H, H, Cin, device="cuda").bfloat16(); w = torch.randn(Cout, 9 * Cin, device="cuda").bfloat16()
out = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.bfloat16)
for _ in range(3): ops.conv3x3(x, w, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.start()
ops.conv3x3(x, w, out=out)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")

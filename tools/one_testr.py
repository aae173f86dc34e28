import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200.init import nondegenerate_init_
from tair_b200.testr import TransformerDetector, default_cfg
dev = torch.device("cuda:0")
det = TransformerDetector(default_cfg("cuda")).to(dev).eval(); nondegenerate_init_(det, 99)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator(device=dev).manual_seed(0)
feats = [torch.randn(s, device=dev, generator=g).bfloat16() for s in ((B, 16, 16, 1280), (B, 32, 32, 1280), (B, 64, 64, 640), (B, 64, 64, 320))]
for _ in range(2): det.testr(feats)
torch.cuda.synchronize()
torch.cuda.profiler.start()
det.testr(feats)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("ok")

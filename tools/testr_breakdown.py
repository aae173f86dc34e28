import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
from tair_b200.init import nondegenerate_init_
from tair_b200.testr import TransformerDetector, default_cfg
dev = torch.device("cuda:0")
det = TransformerDetector(default_cfg("cuda")).to(dev).eval(); nondegenerate_init_(det, 99)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
g = torch.Generator(device=dev).manual_seed(0)
feats = [torch.randn(s, device=dev, generator=g).bfloat16() for s in ((B, 16, 16, 1280), (B, 32, 32, 1280), (B, 64, 64, 640), (B, 64, 64, 320))]
for _ in range(2): det.testr(feats)
torch.cuda.synchronize()
t = ops.KernelTimer(); ops.set_timer(t)
a, b = torch.cuda.Event(True), torch.cuda.Event(True)
a.record(); det.testr(feats); b.record()
ops.set_timer(None); torch.cuda.synchronize()
print("total eager ms", a.elapsed_time(b))
for k, v in sorted(t.summary().items(), key=lambda kv: -kv[1]["ms"]):
    print(f"{k:12s} launches={v['launches']:4d} ms={v['ms']:.3f}")
rows = sorted(t.by_shape().items(), key=lambda kv: -kv[1]["ms"])
for (fam, tag), v in rows[:30]:
    unit = "GB/s" if fam in ("groupnorm", "layernorm", "msda") else "TF/s"
    print(f"{fam:13s} {str(tag):34s} x{v['launches']:3d} {v['ms']:7.3f} ms  {v['work'] / v['ms'] / 1e9:8.0f} {unit} ({v['ms'] / v['launches'] * 1e3:.1f} us each)")

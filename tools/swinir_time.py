import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops
from tair_b200.init import nondegenerate_init_
from tair_b200.model.swinir import SwinIR
m = SwinIR(img_size=64, patch_size=1, in_chans=3, embed_dim=180, depths=[6] * 8, num_heads=[6] * 8, window_size=8, mlp_ratio=2,
           sf=8, img_range=1.0, upsampler="nearest+conv", resi_connection="1conv", unshuffle=True, unshuffle_scale=8).cuda().eval()
nondegenerate_init_(m, 77)
for B in (1, 16):
    x = torch.rand(B, 3, 512, 512, device="cuda")
    for _ in range(3): m(x)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    ops.reset_launch_count()
    a.record()
    for _ in range(5): m(x)
    b.record(); torch.cuda.synchronize()
    print(f"SwinIR B={B}: {a.elapsed_time(b) / 5:.2f} ms per call, {ops.launch_count() // 5} launches, {0.18 * B / (a.elapsed_time(b) / 5e3):.0f} TFLOP/s-equivalent (0.18 TFLOP/tile)")
t = ops.KernelTimer(); ops.set_timer(t); m(x); ops.set_timer(None)
for k, v in sorted(t.summary().items(), key=lambda kv: -kv[1]["ms"]): print(f"  {k:18s} launches={v['launches']:4d} ms={v['ms']:.3f}")

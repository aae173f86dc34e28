"""Bring-up diagnostics for the tcgen05 attention kernel (run on a B200 via gpurun)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.nn.functional as F
from tair_b200 import ops

torch.manual_seed(0)
dev = "cuda"
ok = True


def case(B, H, Lq, Lk, fused=False, scale_in=1.0):
    global ok
    C = H * 64
    if fused and Lq == Lk:
        qkv = (torch.randn(B * Lq, 3 * C, device=dev) * scale_in).bfloat16()
        q, k, v = qkv[:, :C], qkv[:, C:2 * C], qkv[:, 2 * C:]
    else:
        q = (torch.randn(B * Lq, C, device=dev) * scale_in).bfloat16()
        k = (torch.randn(B * Lk, C, device=dev) * scale_in).bfloat16()
        v = torch.randn(B * Lk, C, device=dev).bfloat16()
    out = ops.attention(q, k, v, B=B, H=H, Lq=Lq, Lk=Lk)
    torch.cuda.synchronize()
    qf = q.float().reshape(B, Lq, H, 64).transpose(1, 2)
    kf = k.float().reshape(B, Lk, H, 64).transpose(1, 2)
    vf = v.float().reshape(B, Lk, H, 64).transpose(1, 2)
    ref = F.scaled_dot_product_attention(qf, kf, vf).transpose(1, 2).reshape(B * Lq, C)
    err = (out.float() - ref).abs().max().item() / (ref.abs().max().item() + 1e-9)
    good = err < 2e-2 and bool(torch.isfinite(out.float()).all())
    ok &= good
    print(f"attn B={B} H={H} Lq={Lq} Lk={Lk} fused={fused} s={scale_in} rel_err={err:.3e} {'OK' if good else 'FAIL'}", flush=True)


case(1, 1, 128, 128)
case(1, 1, 128, 256)
case(1, 2, 256, 256, fused=True)
case(2, 5, 1024, 1024, fused=True)
case(2, 5, 4096, 4096, fused=True, scale_in=2.0)
case(2, 10, 1024, 77)
case(2, 20, 256, 77)
case(3, 20, 64, 64, fused=True)
case(3, 20, 64, 77)
case(1, 5, 4096, 77, scale_in=3.0)
case(1, 3, 200, 333)


def bench(B, H, L, Lk=None):
    Lk = Lk or L
    C = H * 64
    q = torch.randn(B * L, C, device=dev).bfloat16()
    k = torch.randn(B * Lk, C, device=dev).bfloat16()
    v = torch.randn(B * Lk, C, device=dev).bfloat16()
    out = torch.empty(B * L, C, device=dev, dtype=torch.bfloat16)
    fl = 4.0 * B * H * L * Lk * 64
    def run(fn, label):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        s, e = torch.cuda.Event(True), torch.cuda.Event(True)
        s.record()
        for _ in range(10):
            fn()
        e.record(); torch.cuda.synchronize()
        ms = s.elapsed_time(e) / 10
        print(f"perf {label} B={B} H={H} L={L} Lk={Lk}: {ms:.3f} ms {fl / ms / 1e9:.1f} TFLOP/s", flush=True)
    run(lambda: ops.attention(q, k, v, B=B, H=H, Lq=L, Lk=Lk, out=out), "tair")
    q4 = q.view(B, L, H, 64).transpose(1, 2); k4 = k.view(B, Lk, H, 64).transpose(1, 2); v4 = v.view(B, Lk, H, 64).transpose(1, 2)
    run(lambda: F.scaled_dot_product_attention(q4, k4, v4), "torch-sdpa")


bench(16, 5, 4096)
bench(16, 10, 1024)
bench(16, 20, 256)
bench(16, 5, 4096, 77)
print("ATTN", "PASS" if ok else "FAIL")
sys.exit(0 if ok else 1)

"""Bring-up diagnostics for the tcgen05 GEMM / implicit conv (run on a B200 via gpurun).

Usage: python tools/gpu_check_gemm.py <group>   (groups: gemm_small gemm_shapes gemm_epi conv conv_s2 perf)
       python tools/gpu_check_gemm.py all       (runs every group in its own subprocess under timeout)
"""
import subprocess
import sys
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

GROUPS = ["gemm_small", "gemm_shapes", "gemm_epi", "conv", "conv_s2", "perf"]


def rel_err(a, b):
    import torch
    a = a.float(); b = b.float()
    return ((a - b).abs().max() / (b.abs().max() + 1e-12)).item()


def run(group):
    import torch
    import torch.nn.functional as F
    from tair_b200 import ops
    torch.manual_seed(0)
    dev = "cuda"
    ok = True

    def gemm_case(M, N, K, **kw):
        nonlocal ok
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        bias = torch.randn(N, device=dev) if kw.get("bias") else None
        act = kw.get("act", 0)
        n_out = N // 2 if act == ops.ACT_GEGLU else N
        res = torch.randn(M, n_out, device=dev).bfloat16() if kw.get("res") else None
        rpg = kw.get("rpg", 0)
        rg = torch.randn((M + rpg - 1) // rpg, N, device=dev) if rpg else None
        out = ops.gemm(a, w, bias=bias, residual=res, rowgroup=rg, rows_per_group=rpg, act=act,
                       out_dtype=kw.get("out_dtype", torch.bfloat16))
        torch.cuda.synchronize()
        ref = a.float() @ w.float().t()
        if bias is not None:
            ref = ref + bias
        if rg is not None:
            ref = ref + rg.repeat_interleave(rpg, 0)[:M]
        if act == ops.ACT_GEGLU:
            bn = 256 if N % 256 == 0 else 128
            r = ref.view(M, N // bn, 2, bn // 2)
            ref = (r[:, :, 0] * F.gelu(r[:, :, 1])).reshape(M, N // 2)
        elif act == ops.ACT_GELU:
            ref = F.gelu(ref)
        elif act == ops.ACT_SILU:
            ref = F.silu(ref)
        elif act == ops.ACT_RELU:
            ref = F.relu(ref)
        if res is not None:
            ref = ref + res.float()
        e = rel_err(out, ref)
        good = e < 2e-2
        ok &= good
        print(f"gemm M={M} N={N} K={K} {kw} rel_err={e:.3e} {'OK' if good else 'FAIL'}", flush=True)

    def conv_case(B, H, W, Cin, Cout, stride=1, **kw):
        nonlocal ok
        x = torch.randn(B, Cin, H, W, device=dev).bfloat16()
        w = (torch.randn(Cout, Cin, 3, 3, device=dev) / (9 * Cin) ** 0.5).bfloat16()
        bias = torch.randn(Cout, device=dev) if kw.get("bias") else None
        xn = x.permute(0, 2, 3, 1).contiguous()
        wp = w.permute(0, 2, 3, 1).reshape(Cout, 9 * Cin).contiguous()
        out = ops.conv3x3(xn, wp, stride=stride, bias=bias)
        torch.cuda.synchronize()
        ref = F.conv2d(x.float(), w.float(), bias, stride=stride, padding=1).permute(0, 2, 3, 1)
        e = rel_err(out, ref)
        good = e < 2e-2
        ok &= good
        print(f"conv B={B} H={H} W={W} Cin={Cin} Cout={Cout} s={stride} {kw} rel_err={e:.3e} {'OK' if good else 'FAIL'}", flush=True)

    if group == "gemm_small":
        gemm_case(128, 128, 64)
        gemm_case(128, 128, 256)
        gemm_case(256, 256, 512)
        gemm_case(128, 64, 64)
        gemm_case(128, 160, 128)
        gemm_case(128, 256, 128)
    elif group == "gemm_shapes":
        gemm_case(4096, 320, 320)
        gemm_case(1000, 320, 320)
        gemm_case(16384, 1280, 640)
        gemm_case(65536, 960, 320)
        gemm_case(1232, 640, 1024)
        gemm_case(16, 1280, 320)
        gemm_case(300, 97, 256)
        gemm_case(300, 4, 320)
        gemm_case(77, 2, 256)
        gemm_case(5000, 1920, 640)
    elif group == "gemm_epi":
        gemm_case(4096, 320, 320, bias=True)
        gemm_case(4096, 320, 320, bias=True, res=True)
        gemm_case(4096, 640, 320, bias=True, rpg=1024)
        gemm_case(4096, 2560, 320, bias=True, act=ops.ACT_GEGLU)
        gemm_case(1000, 2560, 320, bias=True, act=ops.ACT_GEGLU, res=True)
        gemm_case(512, 1024, 256, bias=True, act=ops.ACT_RELU)
        gemm_case(512, 256, 256, bias=True, act=ops.ACT_GELU)
        gemm_case(16, 1280, 320, bias=True, act=ops.ACT_SILU)
        gemm_case(512, 100, 256, bias=True, out_dtype=torch.float32)
        gemm_case(512, 320, 256, bias=True, res=False, out_dtype=torch.float32)
    elif group == "conv":
        conv_case(1, 16, 16, 64, 64)
        conv_case(2, 64, 64, 320, 320, bias=True)
        conv_case(2, 32, 32, 640, 640, bias=True)
        conv_case(2, 16, 16, 1280, 1280)
        conv_case(3, 8, 8, 1280, 1280, bias=True)
        conv_case(1, 8, 8, 2560, 1280)
        conv_case(2, 64, 64, 320, 4, bias=True)
        conv_case(1, 64, 64, 256, 256)
    elif group == "conv_s2":
        conv_case(1, 16, 16, 64, 64, stride=2)
        conv_case(2, 64, 64, 320, 320, stride=2, bias=True)
        conv_case(2, 32, 32, 640, 640, stride=2)
        conv_case(2, 16, 16, 1280, 1280, stride=2)
    elif group == "perf":
        def bench(fn, flops, label):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            s, e2 = torch.cuda.Event(True), torch.cuda.Event(True)
            s.record()
            n = 10
            for _ in range(n):
                fn()
            e2.record()
            torch.cuda.synchronize()
            ms = s.elapsed_time(e2) / n
            print(f"perf {label}: {ms:.3f} ms  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
        for (M, N, K) in [(65536, 320, 320), (65536, 960, 320), (65536, 2560, 320), (16384, 1280, 1280),
                          (16384, 640, 640), (4096, 1280, 1280), (8192, 8192, 8192)]:
            a = torch.randn(M, K, device=dev).bfloat16()
            w = torch.randn(N, K, device=dev).bfloat16()
            out = torch.empty(M, N, device=dev, dtype=torch.bfloat16)
            bench(lambda: ops.gemm(a, w, out=out), 2.0 * M * N * K, f"gemm {M}x{N}x{K}")
            bench(lambda: torch.matmul(a, w.t(), out=out), 2.0 * M * N * K, f"cublas {M}x{N}x{K}")
        for (B, H, Cin, Cout) in [(16, 64, 320, 320), (16, 32, 640, 640), (16, 16, 1280, 1280), (16, 8, 1280, 1280),
                                  (16, 8, 2560, 1280)]:
            x = torch.randn(B, H, H, Cin, device=dev).bfloat16()
            w = torch.randn(Cout, 9 * Cin, device=dev).bfloat16()
            out = torch.empty(B, H, H, Cout, device=dev, dtype=torch.bfloat16)
            bench(lambda: ops.conv3x3(x, w, out=out), 2.0 * B * H * H * Cout * 9 * Cin, f"conv B{B} {H}x{H} {Cin}->{Cout}")
            xc = x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
            wc = w.view(Cout, 3, 3, Cin).permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)
            bench(lambda: F.conv2d(xc, wc, padding=1), 2.0 * B * H * H * Cout * 9 * Cin, f"cudnn B{B} {H}x{H} {Cin}->{Cout}")
    print(f"GROUP {group}: {'PASS' if ok else 'FAIL'}", flush=True)
    return 0 if ok else 1


if __name__ == "__main__":
    g = sys.argv[1] if len(sys.argv) > 1 else "all"
    if g == "all":
        rc = 0
        for grp in GROUPS:
            try:
                r = subprocess.run([sys.executable, __file__, grp], timeout=240)
                rc |= r.returncode
                if r.returncode != 0:
                    print(f"GROUP {grp}: exit {r.returncode}", flush=True)
            except subprocess.TimeoutExpired:
                print(f"GROUP {grp}: TIMEOUT (hang)", flush=True)
                rc |= 1
        sys.exit(rc)
    sys.exit(run(g))

"""A/B of the MSDeformAttn forward on one B200: the UNMODIFIED reference CUDA kernel (ms_deformable_im2col_gpu_kernel,
built from the reference tree into oracle/_ref/libmsda_ref.so by oracle/build_ref.sh) against tair_msda_forward (drop-in
semantics, fp32 and bf16 value) and tair_msda_fused (softmax + location arithmetic + gather in one kernel, what the TESTR
layers call), at the three shapes of one TESTR step (Appendix B of SURVEY.md): encoder Lq 9472, decoder Lq 1600 / 2500,
B = 16 tiles, S = 9472, 8 heads x 32, 4 levels x 4 points.  CUDA-graph replay of 10 back-to-back launches, microseconds
per launch; GB/s on the COMPULSORY bytes (value + locations/weights or projection rows + output, each once)."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tair_b200 import ops

REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libmsda_ref.so")
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(0)
shapes = [(16, 16), (32, 32), (64, 64), (64, 64)]
S = sum(h * w for h, w in shapes)
B, M, D, L, P = int(os.environ.get("B", "16")), 8, 32, 4, 4
shp = torch.tensor(shapes, device=dev, dtype=torch.long)
start = torch.cat([shp.new_zeros(1), (shp[:, 0] * shp[:, 1]).cumsum(0)[:-1]])
ref_lib = C.CDLL(REF) if os.path.exists(REF) else None


def timeit(fn, n=10, reps=5):
    fn(); torch.cuda.synchronize()
    gr = torch.cuda.CUDAGraph()
    with torch.cuda.graph(gr):
        for _ in range(n): fn()
    gr.replay(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(True), torch.cuda.Event(True)
    a.record()
    for _ in range(reps): gr.replay()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps / n * 1e3


out = {}
for name, Lq, ref_dim in (("encoder", 9472, 2), ("decoder_points", 1600, 4), ("decoder_text", 2500, 4)):
    value32 = torch.randn(B, S, M, D, device=dev, generator=g)
    value16 = value32.bfloat16()
    proj = torch.randn(B * Lq, M * L * P * 3, device=dev, generator=g)
    off = proj[:, :M * L * P * 2].view(B, Lq, M, L, P, 2)
    aw = torch.softmax(proj[:, M * L * P * 2:].view(B, Lq, M, L * P), -1).view(B, Lq, M, L, P).contiguous()
    if ref_dim == 2:
        ref = torch.rand(Lq, L, 2, device=dev, generator=g)
        norm = torch.stack([shp[:, 1], shp[:, 0]], -1).float()
        loc = (ref[None, :, None, :, None, :] + off / norm[None, None, None, :, None, :]).contiguous()
        fused = lambda: ops.msda_fused(value16, shp, start, proj_bf, ref, B=B, Lq=Lq, n_heads=M, n_levels=L, n_points=P, ref_shared=True)
    else:
        qpr = 16 if Lq == 1600 else 25
        ref = torch.rand(B, Lq // qpr, L, 4, device=dev, generator=g) * 0.5 + 0.25
        r = ref.repeat_interleave(qpr, 1)
        loc = (r[:, :, None, :, None, :2] + off / P * r[:, :, None, :, None, 2:] * 0.5).contiguous()
        fused = lambda: ops.msda_fused(value16, shp, start, proj_bf, ref, B=B, Lq=Lq, n_heads=M, n_levels=L, n_points=P, q_per_ref=qpr)
    proj_bf = proj.bfloat16()
    res = {}
    o_ours = ops.msda_forward(value32, shp, start, loc, aw)
    if ref_lib is not None:
        o_ref = torch.empty(B, Lq, M * D, device=dev)
        call = lambda: ref_lib.msda_ref_forward_f32(C.c_void_p(value32.data_ptr()), C.c_void_p(shp.data_ptr()), C.c_void_p(start.data_ptr()),
                                                    C.c_void_p(loc.data_ptr()), C.c_void_p(aw.data_ptr()), C.c_void_p(o_ref.data_ptr()),
                                                    B, S, M, D, L, Lq, P, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        call(); torch.cuda.synchronize()
        res["max_abs_diff_vs_reference_kernel_fp32"] = (o_ours - o_ref).abs().max().item()
        res["reference_kernel_fp32_us"] = round(timeit(call), 2)
    res["tair_msda_forward_fp32_us"] = round(timeit(lambda: ops.msda_forward(value32, shp, start, loc, aw)), 2)
    res["tair_msda_forward_bf16_us"] = round(timeit(lambda: ops.msda_forward(value16, shp, start, loc, aw)), 2)
    res["tair_msda_fused_bf16_us"] = round(timeit(fused), 2)
    # compulsory bytes: value + per-query inputs + output, each touched once
    b_ref = value32.numel() * 4 + loc.numel() * 4 + aw.numel() * 4 + B * Lq * M * D * 4
    b_fused = value16.numel() * 2 + proj_bf.numel() * 2 + B * Lq * M * D * 2
    taps = B * Lq * M * L * P * 4 * D     # gathered elements (4 bilinear corners per sample)
    if "reference_kernel_fp32_us" in res:
        res["reference_GBs_compulsory"] = round(b_ref / res["reference_kernel_fp32_us"] / 1e3)
        res["speedup_fused_vs_reference"] = round(res["reference_kernel_fp32_us"] / res["tair_msda_fused_bf16_us"], 2)
        res["speedup_dropin_fp32_vs_reference"] = round(res["reference_kernel_fp32_us"] / res["tair_msda_forward_fp32_us"], 2)
    res["fused_GBs_compulsory"] = round(b_fused / res["tair_msda_fused_bf16_us"] / 1e3)
    res["fused_gather_GBs_from_L1_L2"] = round(taps * 2 / res["tair_msda_fused_bf16_us"] / 1e3)
    out[f"{name}_Lq{Lq}"] = res
print(json.dumps(out, indent=1))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/msda_ab.json", "w"), indent=1)

"""tair_msda_forward against the reference's OWN CUDA kernel (ms_deformable_im2col_gpu_kernel,
testr/adet/layers/csrc/DeformAttn/ms_deform_im2col_cuda.cuh:237-299) compiled unmodified from the reference tree into
oracle/_ref/libmsda_ref.so (oracle/build_ref.sh; the binary travels to the GPU box, the sources do not).  Skipped when the
binary has not been built."""
import ctypes as C
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle", "_ref", "libmsda_ref.so")


@pytest.mark.parametrize("B,Lq,shapes", [(2, 300, [(16, 16), (32, 32), (64, 64), (64, 64)]), (3, 77, [(8, 8), (4, 6), (3, 3)]),
                                         (1, 9472, [(16, 16), (32, 32), (64, 64), (64, 64)])])
def test_drop_in_forward_equals_reference_cuda_kernel(cuda_lib, B, Lq, shapes):
    if not os.path.exists(REF):
        pytest.skip("oracle/_ref/libmsda_ref.so not built (needs the reference tree: sh oracle/build_ref.sh)")
    from tair_b200 import ops
    lib = C.CDLL(REF)
    M, D, P = 8, 32, 4
    L = len(shapes)
    S = sum(h * w for h, w in shapes)
    g = torch.Generator(device="cuda").manual_seed(5)
    value = torch.randn(B, S, M, D, device="cuda", generator=g)
    loc = (torch.rand(B, Lq, M, L, P, 2, device="cuda", generator=g) * 1.4 - 0.2).contiguous()   # some samples fall outside
    w = torch.softmax(torch.randn(B, Lq, M, L * P, device="cuda", generator=g), -1).view(B, Lq, M, L, P).contiguous()
    shp = torch.tensor(shapes, device="cuda", dtype=torch.long)
    start = torch.cat([shp.new_zeros(1), (shp[:, 0] * shp[:, 1]).cumsum(0)[:-1]]).contiguous()
    ref = torch.empty(B, Lq, M * D, device="cuda")
    rc = lib.msda_ref_forward_f32(C.c_void_p(value.data_ptr()), C.c_void_p(shp.data_ptr()), C.c_void_p(start.data_ptr()),
                                  C.c_void_p(loc.data_ptr()), C.c_void_p(w.data_ptr()), C.c_void_p(ref.data_ptr()),
                                  B, S, M, D, L, Lq, P, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    assert rc == 0
    out = ops.msda_forward(value, shp, start, loc, w)
    assert out.shape == ref.shape and (out - ref).abs().max().item() <= 1e-5

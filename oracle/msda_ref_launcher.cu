// TEST / BENCH INFRASTRUCTURE — C-ABI launcher around the UNMODIFIED reference MSDeformAttn forward kernel.
//
// The reference host wrapper (testr/adet/layers/csrc/DeformAttn/ms_deform_attn_cuda.cu:64) no longer compiles against
// torch 2.x (AT_DISPATCH on value.type()), but its kernel header does.  This file #includes that header where it lies
// under the reference tree (never copied into this repository) and exposes ms_deformable_im2col_cuda<float>
// (ms_deform_im2col_cuda.cuh:923-954) so that tools/msda_ab.py can time the reference kernel — the kernel to beat —
// next to tair_msda_forward / tair_msda_fused on the same B200.  Built by oracle/build_ref.sh into oracle/_ref/ (git-ignored,
// travels to the GPU box as a binary).  Nothing under tair_b200/ links or loads it.
#include <cuda_runtime.h>
#include <stdint.h>

#include "ms_deform_im2col_cuda.cuh"

extern "C" int msda_ref_forward_f32(const float* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                    const float* sampling_loc, const float* attn_weight, float* out, int B, int S, int M,
                                    int D, int L, int Lq, int P, void* stream) {
  // the reference wrapper loops over the batch in chunks of im2col_step = min(B, 64) (ms_deform_attn_cuda.cu:50-78);
  // with B <= 64 that is one launch over the whole batch
  ms_deformable_im2col_cuda<float>(static_cast<cudaStream_t>(stream), value, spatial_shapes, level_start_index,
                                   sampling_loc, attn_weight, B, S, M, D, L, Lq, P, out);
  return (int)cudaGetLastError();
}

// Short-sequence multi-head attention for the TESTR composite decoder.
//
// Replaces the nn.MultiheadAttention cores of DeformableCompositeTransformerDecoderLayer
// (testr/adet/layers/deformable_transformer.py:454-466 intra/inter, :485-503 text): sequences of 16 / 25 / 100
// tokens, 8 heads of 32 channels.  These are far too small for the tensor-core flash kernel, so one warp owns
// one (sequence, head): K and V of the head live in shared memory, lanes split the keys for QK^T and the
// channels for PV, softmax is a warp reduction.  Inputs are row-strided views of the fused in_proj output
// [rows, 3*E] so the "inter" (object-wise) attention needs no transpose: a sequence is addressed as
// base = outer*outer_stride + inner*inner_stride, token t at base + t*tok_stride (strides in rows).
#include <atomic>

#include "../../include/tair_b200.h"
#include "common.cuh"

namespace tair {
extern std::atomic<int64_t> g_launch_count;
namespace {

constexpr int MHA_D = 32;
constexpr int MHA_MAX_L = 128;

struct MhaParams {
  const __nv_bfloat16* qkv;  // [rows, ld]: q at col 0, k at col E, v at col 2E (E = H*32)
  __nv_bfloat16* out;        // [rows, ldo]
  int64_t ld, ldo;
  int H, L, E;
  int n_inner;               // sequences = n_outer * n_inner
  int64_t outer_stride, inner_stride, tok_stride;  // in rows
  float scale;
  long n_seq;
};

constexpr int MHA_WARPS = 4;
constexpr int MHA_ROW = MHA_D + 1;  // padded row: lanes reading different keys hit different banks

// block = 4 warps; warp w handles (sequence, head) pair index blockIdx.x*4 + w.
// dynamic smem: per warp K[L][33] then V[L][33] floats
__global__ void __launch_bounds__(MHA_WARPS * 32) mha_small_kernel(const MhaParams p) {
  extern __shared__ float mha_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float (*sk)[MHA_ROW] = reinterpret_cast<float (*)[MHA_ROW]>(mha_smem + (size_t)warp * 2 * p.L * MHA_ROW);
  float (*sv)[MHA_ROW] = sk + p.L;
  const long pair = (long)blockIdx.x * MHA_WARPS + warp;
  if (pair >= p.n_seq * p.H) return;
  const int h = (int)(pair % p.H);
  const long seq = pair / p.H;
  const long base = (seq / p.n_inner) * p.outer_stride + (seq % p.n_inner) * p.inner_stride;
  const __nv_bfloat16* q0 = p.qkv + base * p.ld + h * MHA_D;
  for (int t = 0; t < p.L; ++t) {
    const __nv_bfloat16* r = q0 + (int64_t)t * p.tok_stride * p.ld;
    sk[t][lane] = __bfloat162float(r[p.E + lane]);
    sv[t][lane] = __bfloat162float(r[2 * p.E + lane]);
  }
  __syncwarp();
  for (int i = 0; i < p.L; ++i) {
    const float qv = __bfloat162float(q0[(int64_t)i * p.tok_stride * p.ld + lane]) * p.scale;
    float s[4];
    float mx = -3.0e38f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = lane + c * 32;
      float acc = 0.f;
      if (j < p.L) {
#pragma unroll
        for (int d = 0; d < MHA_D; ++d) acc = fmaf(__shfl_sync(0xffffffffu, qv, d), sk[j][d], acc);
      } else {
#pragma unroll
        for (int d = 0; d < MHA_D; ++d) (void)__shfl_sync(0xffffffffu, qv, d);
        acc = -3.0e38f;
      }
      s[c] = acc;
      mx = fmaxf(mx, acc);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int j = lane + c * 32;
      s[c] = (j < p.L) ? __expf(s[c] - mx) : 0.f;
      sum += s[c];
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    float o_acc = 0.f;  // lane = output channel
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int jmax = p.L - c * 32;
      for (int jj = 0; jj < 32; ++jj) {
        const float pj = __shfl_sync(0xffffffffu, s[c], jj);
        if (jj < jmax) o_acc = fmaf(pj, sv[c * 32 + jj][lane], o_acc);
      }
    }
    p.out[(base + (int64_t)i * p.tok_stride) * p.ldo + h * MHA_D + lane] = __float2bfloat16(o_acc / sum);
  }
}

}  // namespace
}  // namespace tair

using namespace tair;

extern "C" int tair_mha_small(const void* qkv, int64_t ld, void* out, int64_t ldo, int32_t H, int32_t head_dim,
                              int32_t L, int64_t n_outer, int32_t n_inner, int64_t outer_stride, int64_t inner_stride,
                              int64_t tok_stride, float scale, void* stream) {
  TAIR_REQUIRE(qkv && out, "mha_small: NULL pointer");
  TAIR_REQUIRE(head_dim == MHA_D, "mha_small: head_dim must be %d (got %d)", MHA_D, head_dim);
  TAIR_REQUIRE(L > 0 && L <= MHA_MAX_L, "mha_small: sequence length must be in [1, %d] (got %d)", MHA_MAX_L, L);
  TAIR_REQUIRE(H > 0 && n_outer > 0 && n_inner > 0, "mha_small: bad shape");
  TAIR_REQUIRE(ld >= 3 * (int64_t)H * MHA_D && ldo >= (int64_t)H * MHA_D, "mha_small: row stride too small");
  MhaParams p{};
  p.qkv = reinterpret_cast<const __nv_bfloat16*>(qkv);
  p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.ld = ld; p.ldo = ldo; p.H = H; p.L = L; p.E = H * MHA_D; p.n_inner = n_inner;
  p.outer_stride = outer_stride; p.inner_stride = inner_stride; p.tok_stride = tok_stride;
  p.scale = scale; p.n_seq = (long)n_outer * n_inner;
  const long pairs = p.n_seq * H;
  const long grid = (pairs + MHA_WARPS - 1) / MHA_WARPS;
  TAIR_REQUIRE(grid < (1l << 31), "mha_small: problem too large");
  static bool attr = false;
  if (!attr) {
    TAIR_CUDA(cudaFuncSetAttribute(mha_small_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                   MHA_WARPS * 2 * MHA_MAX_L * MHA_ROW * (int)sizeof(float)));
    attr = true;
  }
  const size_t smem = (size_t)MHA_WARPS * 2 * L * MHA_ROW * sizeof(float);
  mha_small_kernel<<<(unsigned)grid, MHA_WARPS * 32, smem, static_cast<cudaStream_t>(stream)>>>(p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("mha_small_kernel");
}

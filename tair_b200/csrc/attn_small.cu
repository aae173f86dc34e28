// Self-attention over many SHORT sequences with 32-wide heads: the nn.MultiheadAttention cores of the TESTR decoder
// (testr/adet/layers/deformable_transformer.py:454-466,485-503: intra-object attention over 16 control points or 25
// characters, inter-object attention over 100 proposals; 8 heads of 32 channels).
//
// Why this is not the tcgen05 flash kernel (attn_tc.cu): a sequence of 16-100 tokens fills 12-78 % of one 128-row UMMA
// tile and a 32-wide head half of its 64-column slot, so that kernel ran these launches at 47-80 us on zero-padded
// [rows, 3*8*64] projections while the work is ~0.1 GFLOP per launch and the data 80 MB: the op is bound by HBM bytes
// and launch structure, not by tensor throughput.  Here one WARP owns 16-query blocks of one (sequence, head), Q / K / V
// head slices (64 B per token) are staged once in shared memory, S = Q K^T and O = P V are register-level
// mma.sync.m16n8k16 (bf16 in, fp32 accumulate; P rounded to bf16 for P V exactly like the flash kernel), and the
// projections stay UNPADDED ([rows, 3*8*32]: half the bytes of the in_proj GEMM output and half the K of out_proj).
//
// Addressing is that of tair_attention_seq_bf16: sequence (o, i) starts at row o*outer_stride + i*inner_stride, token t
// at + t*tok_stride, so the intra / inter "swapdims" of the reference are strides, not copies.
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>

#include "../../include/tair_b200.h"
#include "common.cuh"

namespace tair {
extern std::atomic<int64_t> g_launch_count;
namespace {

constexpr int HD = 32;        // head width
constexpr int LDS_ROW = 40;   // bf16 elements per staged row (80 B): 8 consecutive rows land on disjoint bank groups

__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// four transposed 8x8 bf16 matrices: lanes 8j .. 8j+7 supply the row addresses of matrix j
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

struct SmallParams {
  const __nv_bfloat16 *q, *k, *v;
  __nv_bfloat16* o;
  int64_t ld, ldo, n_items;   // items = sequences x heads
  int64_t outer_stride, inner_stride, tok_stride;
  int H, L, n_inner;
  float scale_log2e;
};

// NKT: key tiles of 8 (padded sequence length LP = 8 * NKT, even NKT).  WPI: warps per item - 1: every warp of the CTA
// owns its own (sequence, head) and all of its query blocks; 4: the CTA owns one item, warps take query blocks w, w+4, ..
template <int NKT, int WPI>
__global__ void __launch_bounds__(128) attn_small_kernel(const SmallParams p) {
  constexpr int LP = NKT * 8;
  constexpr int IPC = 4 / WPI;                       // items per CTA
  constexpr int ITEM_ELEMS = 3 * LP * LDS_ROW;       // Q | K | V staging of one item
  __shared__ __align__(16) __nv_bfloat16 smem[IPC * ITEM_ELEMS];
  pdl_grid_sync();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slot = WPI == 1 ? warp : 0;
  const int64_t item = (int64_t)blockIdx.x * IPC + slot;
  const bool active = item < p.n_items;
  __nv_bfloat16* Qs = smem + slot * ITEM_ELEMS;
  __nv_bfloat16* Ks = Qs + LP * LDS_ROW;
  __nv_bfloat16* Vs = Ks + LP * LDS_ROW;
  int64_t row0 = 0;
  int head = 0;
  if (active) {
    head = (int)(item % p.H);
    const int64_t seq = item / p.H;
    row0 = (seq / p.n_inner) * p.outer_stride + (seq % p.n_inner) * p.inner_stride;
  }
  // ---- stage the head slices: 4 threads x 16 B per token row and matrix; rows >= L are zero-filled ----
  {
    constexpr int NTHR = WPI == 1 ? 32 : 128;
    const int tid = WPI == 1 ? lane : (int)threadIdx.x;
    constexpr int CHUNKS = 3 * LP * 4;
#pragma unroll
    for (int c0 = 0; c0 < CHUNKS; c0 += NTHR) {
      const int c = c0 + tid;
      if (c < CHUNKS) {
        const int mat = c / (LP * 4), rem = c - mat * (LP * 4), tok = rem >> 2, part = rem & 3;
        uint4 val = make_uint4(0, 0, 0, 0);
        if (active && tok < p.L) {
          const __nv_bfloat16* src = (mat == 0 ? p.q : (mat == 1 ? p.k : p.v)) +
                                     (row0 + (int64_t)tok * p.tok_stride) * p.ld + head * HD + part * 8;
          val = __ldg(reinterpret_cast<const uint4*>(src));
        }
        *reinterpret_cast<uint4*>(Qs + (mat * LP + tok) * LDS_ROW + part * 8) = val;
      }
    }
  }
  if (WPI == 1) __syncwarp();
  else __syncthreads();
  if (!active) return;

  const int g = lane >> 2, t = lane & 3;
  const int n_qb = (p.L + 15) >> 4;
  for (int qb = (WPI == 1 ? 0 : warp); qb < n_qb; qb += (WPI == 1 ? 1 : 4)) {
    // A fragments of the 16 x 32 query block (two k-steps of 16 channels)
    uint32_t qa[2][4];
#pragma unroll
    for (int ks = 0; ks < 2; ++ks) {
      const __nv_bfloat16* base = Qs + (qb * 16 + g) * LDS_ROW + ks * 16 + 2 * t;
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(base);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(base + 8 * LDS_ROW);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(base + 8);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(base + 8 * LDS_ROW + 8);
    }
    // S = Q K^T: B fragment (k = channels 2t, 2t+1 [+8]; n = key g) is two 32-bit reads of K's row-major rows
    float s[NKT][4];
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 2; ++ks) {
        const __nv_bfloat16* kb = Ks + (nt * 8 + g) * LDS_ROW + ks * 16 + 2 * t;
        mma_bf16_16816(s[nt], qa[ks], *reinterpret_cast<const uint32_t*>(kb), *reinterpret_cast<const uint32_t*>(kb + 8));
      }
    }
    // softmax over keys; this thread holds rows g (c0, c1) and g + 8 (c2, c3), keys nt*8 + 2t + {0, 1}
    float m0 = -INFINITY, m1 = -INFINITY;
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
      const int key = nt * 8 + 2 * t;
      if (key >= p.L) s[nt][0] = s[nt][2] = -INFINITY;
      if (key + 1 >= p.L) s[nt][1] = s[nt][3] = -INFINITY;
      m0 = fmaxf(m0, fmaxf(s[nt][0], s[nt][1]));
      m1 = fmaxf(m1, fmaxf(s[nt][2], s[nt][3]));
    }
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 1));
    m0 = fmaxf(m0, __shfl_xor_sync(0xffffffffu, m0, 2));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 1));
    m1 = fmaxf(m1, __shfl_xor_sync(0xffffffffu, m1, 2));
    const float off0 = m0 * p.scale_log2e, off1 = m1 * p.scale_log2e;   // key 0 is always valid: finite
    float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
    for (int nt = 0; nt < NKT; ++nt) {
      s[nt][0] = ex2_approx(fmaf(s[nt][0], p.scale_log2e, -off0));
      s[nt][1] = ex2_approx(fmaf(s[nt][1], p.scale_log2e, -off0));
      s[nt][2] = ex2_approx(fmaf(s[nt][2], p.scale_log2e, -off1));
      s[nt][3] = ex2_approx(fmaf(s[nt][3], p.scale_log2e, -off1));
      sum0 += s[nt][0] + s[nt][1];
      sum1 += s[nt][2] + s[nt][3];
    }
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 1);
    sum0 += __shfl_xor_sync(0xffffffffu, sum0, 2);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 1);
    sum1 += __shfl_xor_sync(0xffffffffu, sum1, 2);
    // O = P V: the C fragments of two key tiles are the A fragment of one 16-key step; V^T fragments by ldmatrix.trans
    float o[4][4];
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) o[dn][0] = o[dn][1] = o[dn][2] = o[dn][3] = 0.f;
#pragma unroll
    for (int kt = 0; kt < NKT / 2; ++kt) {
      uint32_t pa[4];
      pa[0] = pack_bf16(s[2 * kt][0], s[2 * kt][1]);
      pa[1] = pack_bf16(s[2 * kt][2], s[2 * kt][3]);
      pa[2] = pack_bf16(s[2 * kt + 1][0], s[2 * kt + 1][1]);
      pa[3] = pack_bf16(s[2 * kt + 1][2], s[2 * kt + 1][3]);
#pragma unroll
      for (int dp = 0; dp < 2; ++dp) {   // channel tiles 2dp, 2dp + 1
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, smem_u32(Vs + (kt * 16 + (lane & 15)) * LDS_ROW + dp * 16 + (lane >> 4) * 8));
        mma_bf16_16816(o[2 * dp], pa, vb[0], vb[1]);
        mma_bf16_16816(o[2 * dp + 1], pa, vb[2], vb[3]);
      }
    }
    const float inv0 = 1.f / sum0, inv1 = 1.f / sum1;
    const int q0 = qb * 16 + g, q1 = q0 + 8;
#pragma unroll
    for (int dn = 0; dn < 4; ++dn) {
      if (q0 < p.L)
        *reinterpret_cast<uint32_t*>(p.o + (row0 + (int64_t)q0 * p.tok_stride) * p.ldo + head * HD + dn * 8 + 2 * t) =
            pack_bf16(o[dn][0] * inv0, o[dn][1] * inv0);
      if (q1 < p.L)
        *reinterpret_cast<uint32_t*>(p.o + (row0 + (int64_t)q1 * p.tok_stride) * p.ldo + head * HD + dn * 8 + 2 * t) =
            pack_bf16(o[dn][2] * inv1, o[dn][3] * inv1);
    }
  }
}

template <int NKT, int WPI>
int launch_small(const SmallParams& p, cudaStream_t st) {
  constexpr int IPC = 4 / WPI;
  const int64_t ctas = (p.n_items + IPC - 1) / IPC;
  TAIR_REQUIRE(ctas < (1ll << 31), "attention_seq32: too many sequences");
  TAIR_LAUNCH((attn_small_kernel<NKT, WPI>), (unsigned)ctas, 128, 0, st, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("attn_small_kernel");
}

}  // namespace
}  // namespace tair

using namespace tair;

extern "C" int tair_attention_seq32_bf16(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo,
                                         int32_t H, int32_t L, int64_t n_outer, int32_t n_inner, int64_t outer_stride,
                                         int64_t inner_stride, int64_t tok_stride, float scale, void* stream) {
  TAIR_REQUIRE(q && k && v && o, "attention_seq32: NULL pointer");
  TAIR_REQUIRE(H > 0 && L > 0 && n_outer > 0 && n_inner > 0 && tok_stride > 0, "attention_seq32: bad shape");
  TAIR_REQUIRE(L <= 128, "attention_seq32: sequences of at most 128 tokens (L=%d); longer ones go to tair_attention_seq_bf16", L);
  TAIR_REQUIRE(ld % 8 == 0 && ldo % 2 == 0 && ld >= H * 32 && ldo >= H * 32, "attention_seq32: bad row strides");
  for (const void* ptr : {q, k, v})
    TAIR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) % 16) == 0, "attention_seq32: q / k / v must be 16-byte aligned");
  TAIR_REQUIRE((reinterpret_cast<uintptr_t>(o) % 4) == 0, "attention_seq32: o must be 4-byte aligned");
  SmallParams p;
  p.q = reinterpret_cast<const __nv_bfloat16*>(q);
  p.k = reinterpret_cast<const __nv_bfloat16*>(k);
  p.v = reinterpret_cast<const __nv_bfloat16*>(v);
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.ld = ld; p.ldo = ldo;
  p.n_items = n_outer * n_inner * H;
  p.outer_stride = outer_stride; p.inner_stride = inner_stride; p.tok_stride = tok_stride;
  p.H = H; p.L = L; p.n_inner = n_inner;
  p.scale_log2e = scale * 1.4426950408889634f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (L <= 16) return launch_small<2, 1>(p, st);
  if (L <= 32) return launch_small<4, 1>(p, st);
  if (L <= 112) return launch_small<14, 4>(p, st);
  return launch_small<16, 4>(p, st);
}

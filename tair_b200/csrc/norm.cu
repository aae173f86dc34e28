// Memory-bound normalisation kernels on channels-last bf16 activations.
//
//   tair_groupnorm_nhwc : GroupNorm(G groups, eps) [+ SiLU | GELU] in two passes
//       reference: GroupNorm32 -> SiLU of every ResBlock (terediff/model/unet.py:148-151,171-178,
//       util.py:182-193, eps 1e-5), Normalize() of SpatialTransformer (attention.py:48-51, eps 1e-6),
//       GroupNorm -> GELU of TESTR's diff_feat_proj (testr/adet/modeling/testr/models.py:76-88).
//   tair_layernorm      : row LayerNorm (attention.py:252-254; deformable_transformer.py norm*).
//
// Layout: x is [B, HW, C] (C contiguous).  Every thread owns a fixed 8-channel vector (16 bytes) and
// walks rows, so loads/stores are 16-byte and fully coalesced; statistics are accumulated in fp32.
// Pass 1 writes per-slab partial sums (no atomics anywhere: bit-reproducible); pass 2 folds them, normalises,
// applies the activation and writes bf16.  Algorithmic traffic: 2 reads + 1 write of the tensor.
#include <atomic>
#include <cstdlib>

#include "../../include/tair_b200.h"
#include "common.cuh"

namespace tair {
extern std::atomic<int64_t> g_launch_count;
namespace {

constexpr int GN_MAX_SLABS = 64;

struct GnParams {
  const __nv_bfloat16* x;
  __nv_bfloat16* y;
  const float* gamma;
  const float* beta;
  float* ws;  // [B, slabs, G, 2]
  int B, HW, C, G, cpg;
  int slabs, rows_per_slab, rows_per_iter, vec_per_row;
  float eps;
  int act;
};

__device__ __forceinline__ void load8(const __nv_bfloat16* p, float (&f)[8]) {
  uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), c = unpack_bf16(q.z), d = unpack_bf16(q.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const float (&f)[8]) {
  uint4 q;
  q.x = pack_bf16(f[0], f[1]); q.y = pack_bf16(f[2], f[3]);
  q.z = pack_bf16(f[4], f[5]); q.w = pack_bf16(f[6], f[7]);
  *reinterpret_cast<uint4*>(p) = q;
}

// grid (slabs, B); block = vec_per_row * rows_per_iter threads (<= 320, so <= 2560 channel slots)
__global__ void gn_stats_kernel(const GnParams p) {
  __shared__ float s_part[2][2560];  // per (row-phase, channel) partial sum / sum of squares
  const int b = blockIdx.y, slab = blockIdx.x;
  const int vec = threadIdx.x % p.vec_per_row;
  const int rsub = threadIdx.x / p.vec_per_row;
  const int c0 = vec * 8;
  // per-channel partial sums (the thread's 8 channels are fixed across rows)
  float sm[8], sq[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { sm[i] = 0.f; sq[i] = 0.f; }
  const int r0 = slab * p.rows_per_slab;
  int r1 = r0 + p.rows_per_slab;
  if (r1 > p.HW) r1 = p.HW;
  const __nv_bfloat16* xb = p.x + (int64_t)b * p.HW * p.C + c0;
  // 4 independent 16-byte loads in flight per thread (memory-level parallelism), then a scalar tail
  int r = r0 + rsub;
  const int step = p.rows_per_iter;
  for (; r + 3 * step < r1; r += 4 * step) {
    uint4 q[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)(r + u * step) * p.C));
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const float2 a = unpack_bf16(q[u].x), b2 = unpack_bf16(q[u].y), c = unpack_bf16(q[u].z), d = unpack_bf16(q[u].w);
      const float f[8] = {a.x, a.y, b2.x, b2.y, c.x, c.y, d.x, d.y};
#pragma unroll
      for (int i = 0; i < 8; ++i) { sm[i] += f[i]; sq[i] = fmaf(f[i], f[i], sq[i]); }
    }
  }
  for (; r < r1; r += step) {
    float f[8];
    load8(xb + (int64_t)r * p.C, f);
#pragma unroll
    for (int i = 0; i < 8; ++i) { sm[i] += f[i]; sq[i] = fmaf(f[i], f[i], sq[i]); }
  }
  // deterministic block reduction (no atomics: results must not depend on scheduling, the sampler is a 50-step
  // recurrence): slot = rsub * C + channel, then one thread per group sums its slots in a fixed order
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s_part[0][rsub * p.C + c0 + i] = sm[i];
    s_part[1][rsub * p.C + c0 + i] = sq[i];
  }
  __syncthreads();
  if (threadIdx.x < p.G) {
    const int g = threadIdx.x;
    float ps = 0.f, pq = 0.f;
    for (int rs = 0; rs < p.rows_per_iter; ++rs) {
      const int base = rs * p.C + g * p.cpg;
      for (int c = 0; c < p.cpg; ++c) {
        ps += s_part[0][base + c];
        pq += s_part[1][base + c];
      }
    }
    float* w = p.ws + ((int64_t)(b * p.slabs + slab) * p.G) * 2;
    w[2 * g] = ps;
    w[2 * g + 1] = pq;
  }
}

template <int ACT>
__global__ void gn_apply_kernel(const GnParams p) {
  __shared__ float s_mean[64], s_rstd[64];
  const int b = blockIdx.y, slab = blockIdx.x;
  if (threadIdx.x < p.G) {
    float s = 0.f, q = 0.f;
    const float* w = p.ws + (int64_t)b * p.slabs * p.G * 2 + threadIdx.x * 2;
    for (int i = 0; i < p.slabs; ++i) {
      s += w[(int64_t)i * p.G * 2];
      q += w[(int64_t)i * p.G * 2 + 1];
    }
    const float n = (float)p.HW * (float)p.cpg;
    const float mean = s / n;
    float var = q / n - mean * mean;
    var = var < 0.f ? 0.f : var;
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + p.eps);
  }
  __syncthreads();
  const int vec = threadIdx.x % p.vec_per_row;
  const int rsub = threadIdx.x / p.vec_per_row;
  const int c0 = vec * 8;
  float sc[8], sh[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int g = (c0 + i) / p.cpg;
    const float ga = __ldg(p.gamma + c0 + i) * s_rstd[g];
    sc[i] = ga;
    sh[i] = __ldg(p.beta + c0 + i) - s_mean[g] * ga;
  }
  const int r0 = slab * p.rows_per_slab;
  int r1 = r0 + p.rows_per_slab;
  if (r1 > p.HW) r1 = p.HW;
  const int64_t base = (int64_t)b * p.HW * p.C + c0;
  auto finish = [&](float (&f)[8], int r) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = fmaf(f[i], sc[i], sh[i]);
      if (ACT == TAIR_ACT_SILU) v = silu_f(v);
      else if (ACT == TAIR_ACT_GELU) v = gelu_f(v);
      f[i] = v;
    }
    store8(p.y + base + (int64_t)r * p.C, f);
  };
  int r = r0 + rsub;
  const int step = p.rows_per_iter;
  for (; r + 3 * step < r1; r += 4 * step) {
    float f[4][8];
#pragma unroll
    for (int u = 0; u < 4; ++u) load8(p.x + base + (int64_t)(r + u * step) * p.C, f[u]);
#pragma unroll
    for (int u = 0; u < 4; ++u) finish(f[u], r + u * step);
  }
  for (; r < r1; r += step) {
    float f[8];
    load8(p.x + base + (int64_t)r * p.C, f);
    finish(f, r);
  }
}

// ---- LayerNorm: one warp per row, values kept in registers, two-pass mean / variance.  (Variants that hoist
// gamma/beta into registers or loop several rows per warp were measured slower on B200: 1.7-1.9 ms vs 1.3 ms per
// denoising step — fewer resident warps to hide the load latency.) ----
template <int MAXV>  // max 16-byte vectors per lane
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y,
                                 int64_t ldy, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 int M, int C, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nvec = C >> 3;
  float v[MAXV][8];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      load8(x + (int64_t)row * ldx + vi * 8, v[k]);
#pragma unroll
      for (int i = 0; i < 8; ++i) s += v[k][i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)C;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float d = v[k][i] - mean;
        q += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)C + eps);
#pragma unroll
  for (int k = 0; k < MAXV; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      float o8[8];
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i) o8[i] = (v[k][i] - mean) * rstd * gg[i] + bb[i];
      store8(y + (int64_t)row * ldy + vi * 8, o8);
    }
  }
}


// LayerNorm over the first Cv channels of rows that are padded to C (C % 8 == 0, C <= 256): the SwinIR trunk keeps its
// 180 channels in 192-wide rows so that every GEMM / conv operand is TMA-legal; pad channels are written as zeros.
__global__ void layernorm_ragged_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y,
                                        int64_t ldy, const float* __restrict__ gamma, const float* __restrict__ beta,
                                        int M, int C, int Cv, float eps) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= M) return;
  const int nvec = C >> 3;
  float v[8];
  float s = 0.f;
  const bool have = lane < nvec;
  if (have) {
    load8(x + (int64_t)row * ldx + lane * 8, v);
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (lane * 8 + i >= Cv) v[i] = 0.f;
      s += v[i];
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  const float mean = s / (float)Cv;
  float q = 0.f;
  if (have) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (lane * 8 + i < Cv) {
        const float d = v[i] - mean;
        q += d * d;
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) q += __shfl_xor_sync(0xffffffffu, q, o);
  const float rstd = rsqrtf(q / (float)Cv + eps);
  if (have) {
    float o8[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = lane * 8 + i;
      o8[i] = c < Cv ? (v[i] - mean) * rstd * __ldg(gamma + c) + __ldg(beta + c) : 0.f;
    }
    store8(y + (int64_t)row * ldy + lane * 8, o8);
  }
}

}  // namespace
}  // namespace tair

using namespace tair;

extern "C" int64_t tair_groupnorm_workspace_bytes(int32_t B, int32_t groups) {
  return (int64_t)B * GN_MAX_SLABS * groups * 2 * (int64_t)sizeof(float);
}

extern "C" int tair_groupnorm_nhwc(const void* x, void* y, const float* gamma, const float* beta, int32_t B,
                                   int32_t HW, int32_t C, int32_t groups, float eps, int32_t act,
                                   void* workspace, void* stream) {
  TAIR_REQUIRE(x && y && gamma && beta && workspace, "groupnorm: NULL pointer");
  TAIR_REQUIRE(B > 0 && HW > 0 && C > 0 && groups > 0 && groups <= 64, "groupnorm: bad shape");
  TAIR_REQUIRE(C % groups == 0 && C % 8 == 0, "groupnorm: C must divide by groups and by 8 (C=%d)", C);
  const int cpg = C / groups;
  TAIR_REQUIRE(act == TAIR_ACT_NONE || act == TAIR_ACT_SILU || act == TAIR_ACT_GELU,
               "groupnorm: activation must be none, SiLU or GELU");
  TAIR_REQUIRE((reinterpret_cast<uintptr_t>(x) % 16) == 0 && (reinterpret_cast<uintptr_t>(y) % 16) == 0,
               "groupnorm: tensors must be 16-byte aligned");
  GnParams p{};
  p.x = reinterpret_cast<const __nv_bfloat16*>(x);
  p.y = reinterpret_cast<__nv_bfloat16*>(y);
  p.gamma = gamma; p.beta = beta; p.ws = reinterpret_cast<float*>(workspace);
  p.B = B; p.HW = HW; p.C = C; p.G = groups; p.cpg = cpg; p.eps = eps; p.act = act;
  p.vec_per_row = C / 8;
  TAIR_REQUIRE(p.vec_per_row <= 1024, "groupnorm: C too large (%d)", C);
  int rpi = 320 / p.vec_per_row;
  if (rpi < 1) rpi = 1;
  if (rpi > HW) rpi = HW;
  p.rows_per_iter = rpi;
  const int threads = p.vec_per_row * rpi;
  // the slab partition depends on the image geometry only (never on the batch size), so a tile's statistics — and
  // with them the whole 50-step trajectory — are bit-identical whatever batch / world size it is processed in
  int slabs = HW / (rpi * 8);  // >= 8 rows per thread
  int max_cfg = 32;   // measured (tools/gn_probe.py): 32 slabs per image beat 64 by ~10 % on the 64x64 / 16x16 levels, 16 lose
  if (const char* e = getenv("TAIR_GN_SLABS")) { const int v = atoi(e); if (v > 0 && v <= GN_MAX_SLABS) max_cfg = v; }   // probe
  if (slabs > max_cfg) slabs = max_cfg;
  const int max_slabs = (HW + rpi - 1) / rpi;
  if (slabs > max_slabs) slabs = max_slabs;
  if (slabs < 1) slabs = 1;
  p.rows_per_slab = (HW + slabs - 1) / slabs;
  p.slabs = (HW + p.rows_per_slab - 1) / p.rows_per_slab;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(p.slabs, B);
  gn_stats_kernel<<<grid, threads, 0, st>>>(p);
  int rc = check_launch("gn_stats_kernel");
  if (rc) return rc;
  if (act == TAIR_ACT_SILU) gn_apply_kernel<TAIR_ACT_SILU><<<grid, threads, 0, st>>>(p);
  else if (act == TAIR_ACT_GELU) gn_apply_kernel<TAIR_ACT_GELU><<<grid, threads, 0, st>>>(p);
  else gn_apply_kernel<TAIR_ACT_NONE><<<grid, threads, 0, st>>>(p);
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  return check_launch("gn_apply_kernel");
}

extern "C" int tair_layernorm(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma,
                              const float* beta, int32_t M, int32_t C, float eps, void* stream) {
  TAIR_REQUIRE(x && y && gamma && beta, "layernorm: NULL pointer");
  TAIR_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && C <= 2048, "layernorm: C must be a multiple of 8 and <= 2048 (C=%d)", C);
  TAIR_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C, "layernorm: bad row strides");
  TAIR_REQUIRE((reinterpret_cast<uintptr_t>(x) % 16) == 0 && (reinterpret_cast<uintptr_t>(y) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(gamma) % 16) == 0 && (reinterpret_cast<uintptr_t>(beta) % 16) == 0,
               "layernorm: tensors must be 16-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int warps = 4;   // tools/ln_probe.py: 4 warps per CTA is as fast as 8 at 320 / 1280 channels and 12 % faster at 640
  if (const char* e = getenv("TAIR_LN_WARPS")) { const int v = atoi(e); if (v >= 1 && v <= 32) warps = v; }   // probe
  const int grid = (M + warps - 1) / warps;
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  __nv_bfloat16* yp = reinterpret_cast<__nv_bfloat16*>(y);
  const int nvec = C / 8;
  if (nvec <= 32) layernorm_kernel<1><<<grid, warps * 32, 0, st>>>(xp, ldx, yp, ldy, gamma, beta, M, C, eps);
  else if (nvec <= 64) layernorm_kernel<2><<<grid, warps * 32, 0, st>>>(xp, ldx, yp, ldy, gamma, beta, M, C, eps);
  else if (nvec <= 160) layernorm_kernel<5><<<grid, warps * 32, 0, st>>>(xp, ldx, yp, ldy, gamma, beta, M, C, eps);
  else layernorm_kernel<8><<<grid, warps * 32, 0, st>>>(xp, ldx, yp, ldy, gamma, beta, M, C, eps);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("layernorm_kernel");
}

extern "C" int tair_layernorm_ragged(const void* x, int64_t ldx, void* y, int64_t ldy, const float* gamma,
                                     const float* beta, int32_t M, int32_t C, int32_t C_valid, float eps, void* stream) {
  TAIR_REQUIRE(x && y && gamma && beta, "layernorm_ragged: NULL pointer");
  TAIR_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && C <= 256 && C_valid > 0 && C_valid <= C,
               "layernorm_ragged: needs C %% 8 == 0, C <= 256, 0 < C_valid <= C (C=%d, C_valid=%d)", C, C_valid);
  TAIR_REQUIRE(ldx % 8 == 0 && ldy % 8 == 0 && ldx >= C && ldy >= C, "layernorm_ragged: bad row strides");
  TAIR_REQUIRE((reinterpret_cast<uintptr_t>(x) % 16) == 0 && (reinterpret_cast<uintptr_t>(y) % 16) == 0,
               "layernorm_ragged: tensors must be 16-byte aligned");
  const int warps = 8;
  layernorm_ragged_kernel<<<(M + warps - 1) / warps, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, gamma, beta, M, C, C_valid, eps);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("layernorm_ragged_kernel");
}

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

// SiLU through the hardware tanh: x * sigmoid(x) = h + h * tanh(h), h = x / 2 — ONE MUFU op (tanh.approx, max relative error
// 2^-11, i.e. below the bf16 rounding of the result) instead of EX2 + RCP.  GroupNorm+SiLU writes 21 M elements per
// launch at the 64x64 level: two MUFU ops per element kept the XU pipe busy for ~10 us of a ~13 us memory-bound kernel.
__device__ __forceinline__ float silu_tanh(float x) {
  const float h = 0.5f * x;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(h));
  return fmaf(h, t, h);
}

// grid (slabs, B); block = vec_per_row * rows_per_iter threads (<= 320, so <= 2560 channel slots).
// U = 16-byte loads in flight per thread: the statistics pass is pure latency hiding (no stores), so every thread issues
// its whole slab share up front whenever it fits.
template <int U>
__global__ void __launch_bounds__(320) gn_stats_kernel(const GnParams p) {
  pdl_grid_sync();
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
  int r = r0 + rsub;
  const int step = p.rows_per_iter;
  for (; r + (U - 1) * step < r1; r += U * step) {
    uint4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(xb + (int64_t)(r + u * step) * p.C));
#pragma unroll
    for (int u = 0; u < U; ++u) {
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
  // recurrence): slot = rsub * C + channel, then one thread per (group, statistic) sums its slots in a fixed order
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s_part[0][rsub * p.C + c0 + i] = sm[i];
    s_part[1][rsub * p.C + c0 + i] = sq[i];
  }
  __syncthreads();
  if (threadIdx.x < 2 * p.G) {
    const int g = threadIdx.x >> 1, which = threadIdx.x & 1;
    const float* sp = s_part[which];
    float acc = 0.f;
    for (int rs = 0; rs < p.rows_per_iter; ++rs) {
      const int base = rs * p.C + g * p.cpg;
      for (int c = 0; c < p.cpg; ++c) acc += sp[base + c];
    }
    p.ws[((int64_t)(b * p.slabs + slab) * p.G) * 2 + threadIdx.x] = acc;
  }
}

template <int ACT, int U>
__global__ void __launch_bounds__(320) gn_apply_kernel(const GnParams p) {
  pdl_grid_sync();
  __shared__ float s_mean[64], s_rstd[64];
  const int b = blockIdx.y, slab = blockIdx.x;
  if (threadIdx.x < p.G) {
    float s = 0.f, q = 0.f;
    const float* w = p.ws + (int64_t)b * p.slabs * p.G * 2 + threadIdx.x * 2;
    for (int i = 0; i < p.slabs; ++i) {
      const float2 t = *reinterpret_cast<const float2*>(w + (int64_t)i * p.G * 2);
      s += t.x;
      q += t.y;
    }
    const float n = (float)p.HW * (float)p.cpg;
    const float mean = s / n;
    float var = q / n - mean * mean;
    var = var < 0.f ? 0.f : var;
    s_mean[threadIdx.x] = mean;
    s_rstd[threadIdx.x] = rsqrtf(var + p.eps);
  }
  const int vec = threadIdx.x % p.vec_per_row;
  const int rsub = threadIdx.x / p.vec_per_row;
  const int c0 = vec * 8;
  const int r0 = slab * p.rows_per_slab;
  int r1 = r0 + p.rows_per_slab;
  if (r1 > p.HW) r1 = p.HW;
  const int64_t base = (int64_t)b * p.HW * p.C + c0;
  const int step = p.rows_per_iter;
  int r = r0 + rsub;
  // the first batch of loads is issued BEFORE the statistics are folded: it does not depend on them
  uint4 q0[U];
  const bool first = r + (U - 1) * step < r1;
  if (first) {
#pragma unroll
    for (int u = 0; u < U; ++u) q0[u] = __ldg(reinterpret_cast<const uint4*>(p.x + base + (int64_t)(r + u * step) * p.C));
  }
  const float4 g0 = __ldg(reinterpret_cast<const float4*>(p.gamma + c0)), g1 = __ldg(reinterpret_cast<const float4*>(p.gamma + c0 + 4));
  const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.beta + c0)), b1 = __ldg(reinterpret_cast<const float4*>(p.beta + c0 + 4));
  __syncthreads();
  float sc[8], sh[8];
  {
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int g = (c0 + i) / p.cpg;
      const float ga = gg[i] * s_rstd[g];
      sc[i] = ga;
      sh[i] = bb[i] - s_mean[g] * ga;
    }
  }
  auto finish = [&](const uint4& q, int row) {
    const float2 a = unpack_bf16(q.x), b2 = unpack_bf16(q.y), c = unpack_bf16(q.z), d = unpack_bf16(q.w);
    float f[8] = {a.x, a.y, b2.x, b2.y, c.x, c.y, d.x, d.y};
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      float v = fmaf(f[i], sc[i], sh[i]);
      if (ACT == TAIR_ACT_SILU) v = silu_tanh(v);
      else if (ACT == TAIR_ACT_GELU) v = gelu_f(v);
      f[i] = v;
    }
    store8(p.y + base + (int64_t)row * p.C, f);
  };
  if (first) {
#pragma unroll
    for (int u = 0; u < U; ++u) finish(q0[u], r + u * step);
    r += U * step;
  }
  for (; r + (U - 1) * step < r1; r += U * step) {
    uint4 q[U];
#pragma unroll
    for (int u = 0; u < U; ++u) q[u] = __ldg(reinterpret_cast<const uint4*>(p.x + base + (int64_t)(r + u * step) * p.C));
#pragma unroll
    for (int u = 0; u < U; ++u) finish(q[u], r + u * step);
  }
  for (; r < r1; r += step) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p.x + base + (int64_t)r * p.C));
    finish(q, r);
  }
}

// Per-row LayerNorm statistics only: out[m] = (mean, rstd) of x[m, :C].  The normalisation itself is folded into the
// GEMM that consumes the LayerNorm output (tair_epilogue.ln_row_stats / ln_col_sum): the activation is read once
// (2 B per element) and never re-written.  One warp per row, values in registers, two-pass mean / variance.
// Mapping: LPR = 2^lpr_log2 lanes share a row (lane `sub` owns vectors sub, sub + LPR, ...: every load instruction covers
// LPR * 16 contiguous bytes per row), a warp covers 32 / LPR row groups x R2 rows each, and ALL loads of a thread
// (VPL * R2 <= 10) are issued before the first use — the kernel is a pure read stream, its speed is bytes in flight.
template <int VPL, int R2>
__global__ void row_stats_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, float2* __restrict__ out, int M, int C,
                                 float eps, int lpr_log2) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int lpr = 1 << lpr_log2;
  const int sub = lane & (lpr - 1), rgrp = lane >> lpr_log2;
  const int rows_per_warp = (32 >> lpr_log2) * R2;
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * rows_per_warp + rgrp * R2;
  const int nvec = C >> 3;
  uint4 raw[R2][VPL];
#pragma unroll
  for (int rr = 0; rr < R2; ++rr) {
    const int row = row0 + rr < M ? row0 + rr : M - 1;   // clamp: rows past the end are loaded but not written
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = sub + k * lpr;
      raw[rr][k] = vi < nvec ? __ldg(reinterpret_cast<const uint4*>(x + (int64_t)row * ldx + vi * 8)) : make_uint4(0, 0, 0, 0);
    }
  }
  float s[R2], q[R2];
#pragma unroll
  for (int rr = 0; rr < R2; ++rr) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const float2 a = unpack_bf16(raw[rr][k].x), b2 = unpack_bf16(raw[rr][k].y), c = unpack_bf16(raw[rr][k].z), d = unpack_bf16(raw[rr][k].w);
      acc += ((a.x + a.y) + (b2.x + b2.y)) + ((c.x + c.y) + (d.x + d.y));   // padding vectors are zero
    }
    s[rr] = acc;
  }
  for (int o = lpr >> 1; o > 0; o >>= 1)
#pragma unroll
    for (int rr = 0; rr < R2; ++rr) s[rr] += __shfl_xor_sync(0xffffffffu, s[rr], o);
#pragma unroll
  for (int rr = 0; rr < R2; ++rr) {
    const float mean = s[rr] / (float)C;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      if (sub + k * lpr < nvec) {
        const float2 a = unpack_bf16(raw[rr][k].x), b2 = unpack_bf16(raw[rr][k].y), c = unpack_bf16(raw[rr][k].z), d = unpack_bf16(raw[rr][k].w);
        const float f[8] = {a.x, a.y, b2.x, b2.y, c.x, c.y, d.x, d.y};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float dd = f[i] - mean;
          acc = fmaf(dd, dd, acc);
        }
      }
    }
    q[rr] = acc;
    s[rr] = mean;
  }
  for (int o = lpr >> 1; o > 0; o >>= 1)
#pragma unroll
    for (int rr = 0; rr < R2; ++rr) q[rr] += __shfl_xor_sync(0xffffffffu, q[rr], o);
  if (sub == 0) {
#pragma unroll
    for (int rr = 0; rr < R2; ++rr)
      if (row0 + rr < M) out[row0 + rr] = make_float2(s[rr], rsqrtf(q[rr] / (float)C + eps));
  }
}

// ---- LayerNorm: one warp per row, values kept in registers, two-pass mean / variance.  (Variants that hoist
// gamma/beta into registers or loop several rows per warp were measured slower on B200: 1.7-1.9 ms vs 1.3 ms per
// denoising step — fewer resident warps to hide the load latency.) ----
template <int MAXV>  // max 16-byte vectors per lane
__global__ void layernorm_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y,
                                 int64_t ldy, const float* __restrict__ gamma, const float* __restrict__ beta,
                                 int M, int C, float eps) {
  pdl_grid_sync();
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


// LayerNorm for narrow rows (C <= 512: TESTR's 256-wide tokens, 151552 rows per encoder layer): one warp per R2 rows with
// ALL of a thread's loads issued before the first use, like row_stats_kernel.  With one row per warp a thread has a
// single 16-byte load in flight and the kernel sits at half the HBM rate (45 us for 151552 x 256, floor 24).
template <int VPL, int R2>
__global__ void layernorm_rows_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y,
                                      int64_t ldy, const float* __restrict__ gamma, const float* __restrict__ beta,
                                      int M, int C, float eps) {
  pdl_grid_sync();
  const int lane = threadIdx.x & 31;
  const int row0 = (blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * R2;
  if (row0 >= M) return;
  const int nvec = C >> 3;
  uint4 raw[R2][VPL];
#pragma unroll
  for (int rr = 0; rr < R2; ++rr) {
    const int row = row0 + rr < M ? row0 + rr : M - 1;   // clamp: rows past the end are loaded but not written
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const int vi = lane + k * 32;
      raw[rr][k] = vi < nvec ? __ldg(reinterpret_cast<const uint4*>(x + (int64_t)row * ldx + vi * 8)) : make_uint4(0, 0, 0, 0);
    }
  }
  // unpacked ONCE and kept in registers (ncu: the first version, which unpacked the bf16 words in each of its three
  // passes, had its issue slots 72 % busy at 43 % of the DRAM rate); the second pass overwrites f with f - mean, which
  // is also what the output pass needs
  float f[R2][VPL][8];
  float s[R2], q[R2];
#pragma unroll
  for (int rr = 0; rr < R2; ++rr) {
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const float2 a = unpack_bf16(raw[rr][k].x), b2 = unpack_bf16(raw[rr][k].y), c = unpack_bf16(raw[rr][k].z), d = unpack_bf16(raw[rr][k].w);
      f[rr][k][0] = a.x; f[rr][k][1] = a.y; f[rr][k][2] = b2.x; f[rr][k][3] = b2.y;
      f[rr][k][4] = c.x; f[rr][k][5] = c.y; f[rr][k][6] = d.x; f[rr][k][7] = d.y;
      acc += ((a.x + a.y) + (b2.x + b2.y)) + ((c.x + c.y) + (d.x + d.y));
    }
    s[rr] = acc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int rr = 0; rr < R2; ++rr) s[rr] += __shfl_xor_sync(0xffffffffu, s[rr], o);
#pragma unroll
  for (int rr = 0; rr < R2; ++rr) {
    const float mean = s[rr] / (float)C;
    float acc = 0.f;
#pragma unroll
    for (int k = 0; k < VPL; ++k) {
      const bool live = lane + k * 32 < nvec;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float dd = f[rr][k][i] - mean;
        f[rr][k][i] = dd;
        if (live) acc = fmaf(dd, dd, acc);
      }
    }
    q[rr] = acc;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1)
#pragma unroll
    for (int rr = 0; rr < R2; ++rr) q[rr] += __shfl_xor_sync(0xffffffffu, q[rr], o);
#pragma unroll
  for (int k = 0; k < VPL; ++k) {
    const int vi = lane + k * 32;
    if (vi < nvec) {
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + vi * 8 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + vi * 8 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int rr = 0; rr < R2; ++rr) {
        if (row0 + rr < M) {
          const float rstd = rsqrtf(q[rr] / (float)C + eps);
          float o8[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) o8[i] = f[rr][k][i] * rstd * gg[i] + bb[i];   // same expression as layernorm_kernel
          store8(y + (int64_t)(row0 + rr) * ldy + vi * 8, o8);
        }
      }
    }
  }
}

// LayerNorm over the first Cv channels of rows that are padded to C (C % 8 == 0, C <= 256): the SwinIR trunk keeps its
// 180 channels in 192-wide rows so that every GEMM / conv operand is TMA-legal; pad channels are written as zeros.
__global__ void layernorm_ragged_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y,
                                        int64_t ldy, const float* __restrict__ gamma, const float* __restrict__ beta,
                                        int M, int C, int Cv, float eps) {
  pdl_grid_sync();
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
  TAIR_REQUIRE(p.vec_per_row <= 320, "groupnorm: C too large (%d > 2560)", C);
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
  // loads in flight per thread: the whole slab share when it is at most 16 rows per thread (the 64x64 / 32x32 levels)
  const int rows_per_thread = (p.rows_per_slab + rpi - 1) / rpi;
  static int u_cfg = -1;
  if (u_cfg < 0) { const char* e = getenv("TAIR_GN_UNROLL"); u_cfg = e ? atoi(e) : 0; }   // probe
  const int U = u_cfg > 0 ? u_cfg : (rows_per_thread >= 16 ? 16 : (rows_per_thread >= 8 ? 8 : (rows_per_thread >= 4 ? 4 : 1)));
  if (U >= 16) TAIR_LAUNCH((gn_stats_kernel<16>), grid, threads, 0, st, p);
  else if (U >= 8) TAIR_LAUNCH((gn_stats_kernel<8>), grid, threads, 0, st, p);
  else if (U >= 4) TAIR_LAUNCH((gn_stats_kernel<4>), grid, threads, 0, st, p);
  else TAIR_LAUNCH((gn_stats_kernel<1>), grid, threads, 0, st, p);
  int rc = check_launch("gn_stats_kernel");
  if (rc) return rc;
  const int UA = U >= 8 ? 8 : (U >= 4 ? 4 : 1);
#define TAIR_GN_APPLY(ACT_)                                                                   \
  do {                                                                                        \
    if (UA == 8) TAIR_LAUNCH((gn_apply_kernel<ACT_, 8>), grid, threads, 0, st, p);            \
    else if (UA == 4) TAIR_LAUNCH((gn_apply_kernel<ACT_, 4>), grid, threads, 0, st, p);       \
    else TAIR_LAUNCH((gn_apply_kernel<ACT_, 1>), grid, threads, 0, st, p);                    \
  } while (0)
  if (act == TAIR_ACT_SILU) TAIR_GN_APPLY(TAIR_ACT_SILU);
  else if (act == TAIR_ACT_GELU) TAIR_GN_APPLY(TAIR_ACT_GELU);
  else TAIR_GN_APPLY(TAIR_ACT_NONE);
#undef TAIR_GN_APPLY
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  return check_launch("gn_apply_kernel");
}

extern "C" int tair_row_stats(const void* x, int64_t ldx, float* out, int32_t M, int32_t C, float eps, void* stream) {
  TAIR_REQUIRE(x && out, "row_stats: NULL pointer");
  TAIR_REQUIRE(M > 0 && C > 0 && C % 8 == 0 && C <= 2048, "row_stats: C must be a multiple of 8 and <= 2048 (C=%d)", C);
  TAIR_REQUIRE(ldx % 8 == 0 && ldx >= C, "row_stats: bad row stride");
  TAIR_REQUIRE((reinterpret_cast<uintptr_t>(x) % 16) == 0 && (reinterpret_cast<uintptr_t>(out) % 8) == 0,
               "row_stats: x must be 16-byte and out 8-byte aligned");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  static int warps_cfg = -1;
  if (warps_cfg < 0) { const char* e = getenv("TAIR_RS_WARPS"); warps_cfg = e ? atoi(e) : 8; }   // probe
  const int warps = warps_cfg >= 1 && warps_cfg <= 32 ? warps_cfg : 8;
  const __nv_bfloat16* xp = reinterpret_cast<const __nv_bfloat16*>(x);
  float2* op = reinterpret_cast<float2*>(out);
  const int nvec = C / 8;
  // lanes per row: the smallest power of two >= nvec / 5 (8 lanes for 320 channels, 16 for 640, 32 for 1280), at least 8
  int lpr_log2 = 3;
  while ((nvec + (1 << lpr_log2) - 1) / (1 << lpr_log2) > 5 && lpr_log2 < 5) ++lpr_log2;
  const int vpl = (nvec + (1 << lpr_log2) - 1) / (1 << lpr_log2);
  auto grid_for = [&](int R2) { const int rows_per_cta = warps * (32 >> lpr_log2) * R2; return (M + rows_per_cta - 1) / rows_per_cta; };
  if (vpl <= 1) TAIR_LAUNCH((row_stats_kernel<1, 4>), grid_for(4), warps * 32, 0, st, xp, ldx, op, M, C, eps, lpr_log2);
  else if (vpl <= 2) TAIR_LAUNCH((row_stats_kernel<2, 4>), grid_for(4), warps * 32, 0, st, xp, ldx, op, M, C, eps, lpr_log2);
  else if (vpl <= 3) TAIR_LAUNCH((row_stats_kernel<3, 2>), grid_for(2), warps * 32, 0, st, xp, ldx, op, M, C, eps, lpr_log2);
  else if (vpl <= 5) TAIR_LAUNCH((row_stats_kernel<5, 2>), grid_for(2), warps * 32, 0, st, xp, ldx, op, M, C, eps, lpr_log2);
  else TAIR_LAUNCH((row_stats_kernel<8, 1>), grid_for(1), warps * 32, 0, st, xp, ldx, op, M, C, eps, lpr_log2);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("row_stats_kernel");
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
  static int rows_mode = -1;   // TAIR_LN_ROWS=0: the one-row-per-warp kernel for every width (A/B probe)
  if (rows_mode < 0) { const char* e = getenv("TAIR_LN_ROWS"); rows_mode = e ? atoi(e) : 1; }
  if (rows_mode && nvec <= 32) {
    TAIR_LAUNCH((layernorm_rows_kernel<1, 4>), (M + warps * 4 - 1) / (warps * 4), warps * 32, 0, st, xp, ldx, yp, ldy, gamma, beta, M, C, eps);
  } else if (rows_mode && nvec <= 64) {
    TAIR_LAUNCH((layernorm_rows_kernel<2, 2>), (M + warps * 2 - 1) / (warps * 2), warps * 32, 0, st, xp, ldx, yp, ldy, gamma, beta, M, C, eps);
  } else
  if (nvec <= 32) TAIR_LAUNCH((layernorm_kernel<1>), grid, warps * 32, 0, st, xp, ldx, yp, ldy, gamma, beta, M, C, eps);
  else if (nvec <= 64) TAIR_LAUNCH((layernorm_kernel<2>), grid, warps * 32, 0, st, xp, ldx, yp, ldy, gamma, beta, M, C, eps);
  else if (nvec <= 160) TAIR_LAUNCH((layernorm_kernel<5>), grid, warps * 32, 0, st, xp, ldx, yp, ldy, gamma, beta, M, C, eps);
  else TAIR_LAUNCH((layernorm_kernel<8>), grid, warps * 32, 0, st, xp, ldx, yp, ldy, gamma, beta, M, C, eps);
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
  TAIR_LAUNCH((layernorm_ragged_kernel), (M + warps - 1) / warps, warps * 32, 0, static_cast<cudaStream_t>(stream), 
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<__nv_bfloat16*>(y), ldy, gamma, beta, M, C, C_valid, eps);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("layernorm_ragged_kernel");
}

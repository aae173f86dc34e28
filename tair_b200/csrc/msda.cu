// Multi-scale deformable attention forward (bilinear gather + weighted sum) for sm_100a.
//
// Drop-in for `_C.ms_deform_attn_forward` (testr/adet/layers/csrc/vision.cpp:52-55 ->
// ms_deform_attn_cuda.cu:20-80 -> ms_deformable_im2col_gpu_kernel, ms_deform_im2col_cuda.cuh:237-299,
// bilinear tap rule :33-84).  Semantics kept exactly:
//   w_im = loc_x * W_l - 0.5, h_im = loc_y * H_l - 0.5; a sample contributes only if
//   h_im > -1 && w_im > -1 && h_im < H_l && w_im < W_l; each of the 4 taps is individually
//   bounds-checked and reads zero outside; value is addressed as ((b*S + start_l + y*W_l + x)*M + m)*D + c.
//
// Mapping (warp-cooperative): one work item = (b, q, head); D/8 lanes share an item, each lane owning 8
// consecutive channels, so every tap is a single 16-byte (bf16) / 32-byte (fp32) load per lane, the lanes
// of a head read one contiguous D-channel segment, and a warp writes 32*8 consecutive output channels.
// The reference kernel uses one thread per output scalar, re-reading every location/weight/shape scalar
// D times and doing the index arithmetic in int64 per tap.
#include <atomic>
#include <cstdlib>

#include "../../include/tair_b200.h"
#include "common.cuh"

namespace tair {
extern std::atomic<int64_t> g_launch_count;
namespace {

struct MsdaParams {
  const void* value;
  const int64_t* shapes;  // [L,2] (H,W), device
  const int64_t* starts;  // [L], device
  const float* loc;       // [B,Lq,M,L,P,2]
  const float* attw;      // [B,Lq,M,L,P]
  void* out;              // [B,Lq,M*D]
  int B, S, M, D, L, Lq, P;
  int lanes_per_item;     // D / 8
  long items;             // B*Lq*M
};

template <typename VT>
__device__ __forceinline__ void load_vec8(const VT* p, float (&f)[8]);
template <>
__device__ __forceinline__ void load_vec8<float>(const float* p, float (&f)[8]) {
  const float4 a = __ldg(reinterpret_cast<const float4*>(p));
  const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w; f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}
template <>
__device__ __forceinline__ void load_vec8<__nv_bfloat16>(const __nv_bfloat16* p, float (&f)[8]) {
  const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
  const float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), c = unpack_bf16(q.z), d = unpack_bf16(q.w);
  f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
}

template <typename VT, typename OT>
__global__ void __launch_bounds__(256) msda_fwd_kernel(const MsdaParams p) {
  pdl_grid_sync();
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long item = gtid / p.lanes_per_item;
  const int part = (int)(gtid - item * p.lanes_per_item);
  if (item >= p.items) return;
  const int m = (int)(item % p.M);
  const long bq = item / p.M;
  const int b = (int)(bq / p.Lq);
  const VT* vbase = reinterpret_cast<const VT*>(p.value) + ((long)b * p.S * p.M + m) * p.D + part * 8;
  const long row_stride = (long)p.M * p.D;
  const float* locp = p.loc + item * p.L * p.P * 2;
  const float* wp = p.attw + item * p.L * p.P;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;

  for (int l = 0; l < p.L; ++l) {
    const int H = (int)__ldg(p.shapes + 2 * l), W = (int)__ldg(p.shapes + 2 * l + 1);
    const VT* vl = vbase + (long)__ldg(p.starts + l) * row_stride;
    for (int s = 0; s < p.P; ++s) {
      const float2 xy = __ldg(reinterpret_cast<const float2*>(locp) + l * p.P + s);
      const float aw = __ldg(wp + l * p.P + s);
      const float h_im = xy.y * (float)H - 0.5f;
      const float w_im = xy.x * (float)W - 0.5f;
      if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
        const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
        const int h_high = h_low + 1, w_high = w_low + 1;
        const float lh = h_im - (float)h_low, lw = w_im - (float)w_low;
        const float hh = 1.f - lh, hw = 1.f - lw;
        float v1[8], v2[8], v3[8], v4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { v1[i] = 0.f; v2[i] = 0.f; v3[i] = 0.f; v4[i] = 0.f; }
        const bool top = h_low >= 0, bot = h_high <= H - 1, left = w_low >= 0, right = w_high <= W - 1;
        if (top && left) load_vec8<VT>(vl + ((long)h_low * W + w_low) * row_stride, v1);
        if (top && right) load_vec8<VT>(vl + ((long)h_low * W + w_high) * row_stride, v2);
        if (bot && left) load_vec8<VT>(vl + ((long)h_high * W + w_low) * row_stride, v3);
        if (bot && right) load_vec8<VT>(vl + ((long)h_high * W + w_high) * row_stride, v4);
        const float w1 = hh * hw, w2 = hh * lw, w3 = lh * hw, w4 = lh * lw;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float val = w1 * v1[i] + w2 * v2[i] + w3 * v3[i] + w4 * v4[i];
          acc[i] += val * aw;
        }
      }
    }
  }
  OT* op = reinterpret_cast<OT*>(p.out) + item * p.D + part * 8;
  if constexpr (sizeof(OT) == 4) {
    reinterpret_cast<float4*>(op)[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
    reinterpret_cast<float4*>(op)[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
  } else {
    uint4 q;
    q.x = pack_bf16(acc[0], acc[1]); q.y = pack_bf16(acc[2], acc[3]);
    q.z = pack_bf16(acc[4], acc[5]); q.w = pack_bf16(acc[6], acc[7]);
    *reinterpret_cast<uint4*>(op) = q;
  }
}


// ---- fused variant: softmax over the L*P logits and the sampling-location arithmetic of
// MSDeformAttn.forward (testr/adet/layers/ms_deform_attn.py:136-149) are done in the kernel, straight from the
// fused [sampling_offsets | attention_weights] projection row, so no location / weight tensor is materialised.
struct MsdaFusedParams {
  const __nv_bfloat16* value;  // [B,S,M,D]
  const int64_t* shapes;
  const int64_t* starts;
  const void* proj;            // [B*Lq, ldp] fp32 or bf16: M*L*P*2 offsets then M*L*P logits
  int64_t ldp;
  const float* ref;            // [ (B) , Lq/q_per_ref, L, ref_dim ]
  int64_t ref_batch_stride;    // elements; 0 = reference points shared by all images
  int ref_dim, q_per_ref;
  __nv_bfloat16* out;          // [B*Lq, M*D]
  int B, S, M, D, L, Lq, P;
  int lanes_per_item;
  int wide_loads;              // 1: 32-byte (LDG.256) corner loads are legal for this launch
  long items;
};

constexpr int MSDA_MAX_LP = 32;

template <int LT, int PT, typename PJ>  // compile-time (levels, points) or 0 for runtime loops; PJ = proj element type
__global__ void __launch_bounds__(256) msda_fused_kernel(const MsdaFusedParams p) {
  pdl_grid_sync();
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long item = gtid / p.lanes_per_item;
  const int part = (int)(gtid - item * p.lanes_per_item);
  if (item >= p.items) return;
  const int m = (int)(item % p.M);
  const long bq = item / p.M;
  const int b = (int)(bq / p.Lq);
  const int q = (int)(bq - (long)b * p.Lq);
  const int nL = LT > 0 ? LT : p.L, nP = PT > 0 ? PT : p.P;
  const int LP = nL * nP;
  const PJ* row = reinterpret_cast<const PJ*>(p.proj) + bq * p.ldp;
  const PJ* offp = row + (long)m * LP * 2;
  const PJ* logit = row + (long)p.M * LP * 2 + (long)m * LP;
  auto ldf = [](const PJ* q) -> float {
    if constexpr (sizeof(PJ) == 4) return __ldg(reinterpret_cast<const float*>(q));
    else return __bfloat162float(__ldg(reinterpret_cast<const __nv_bfloat16*>(q)));
  };
  // softmax over the L*P logits of this (query, head)
  float w[LT > 0 ? LT * PT : MSDA_MAX_LP];
  float mx = -3.0e38f;
#pragma unroll
  for (int i = 0; i < (LT > 0 ? LT * PT : MSDA_MAX_LP); ++i)
    if (i < LP) { w[i] = ldf(logit + i); mx = fmaxf(mx, w[i]); }
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < (LT > 0 ? LT * PT : MSDA_MAX_LP); ++i)
    if (i < LP) { w[i] = __expf(w[i] - mx); sum += w[i]; }
  const float inv = 1.f / sum;
  const float* refq = p.ref + (long)b * p.ref_batch_stride + (long)(q / p.q_per_ref) * p.L * p.ref_dim;
  const __nv_bfloat16* vbase = p.value + ((long)b * p.S * p.M + m) * p.D + part * 8;
  const long row_stride = (long)p.M * p.D;
  float acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = 0.f;
#pragma unroll
  for (int l = 0; l < (LT > 0 ? LT : 8); ++l) {
    if (l >= nL) break;
    const int H = (int)__ldg(p.shapes + 2 * l), W = (int)__ldg(p.shapes + 2 * l + 1);
    const __nv_bfloat16* vl = vbase + (long)__ldg(p.starts + l) * row_stride;
    const float rx = __ldg(refq + l * p.ref_dim), ry = __ldg(refq + l * p.ref_dim + 1);
    float sx, sy;  // offset scale: 1/(W,H) for points, box_wh * 0.5 / P for boxes (ms_deform_attn.py:139-146)
    if (p.ref_dim == 2) { sx = 1.f / (float)W; sy = 1.f / (float)H; }
    else { sx = __ldg(refq + l * p.ref_dim + 2) * 0.5f / (float)nP; sy = __ldg(refq + l * p.ref_dim + 3) * 0.5f / (float)nP; }
#pragma unroll
    for (int s = 0; s < (PT > 0 ? PT : 8); ++s) {
      if (s >= nP) break;
      const float2 off = make_float2(ldf(offp + 2 * (l * nP + s)), ldf(offp + 2 * (l * nP + s) + 1));
      const float aw = w[l * nP + s] * inv;
      const float h_im = (ry + off.y * sy) * (float)H - 0.5f;
      const float w_im = (rx + off.x * sx) * (float)W - 0.5f;
      if (h_im > -1.f && w_im > -1.f && h_im < (float)H && w_im < (float)W) {
        const int h_low = (int)floorf(h_im), w_low = (int)floorf(w_im);
        const int h_high = h_low + 1, w_high = w_low + 1;
        const float lh = h_im - (float)h_low, lw = w_im - (float)w_low;
        const float hh = 1.f - lh, hw = 1.f - lw;
        float v1[8], v2[8], v3[8], v4[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) { v1[i] = 0.f; v2[i] = 0.f; v3[i] = 0.f; v4[i] = 0.f; }
        const bool top = h_low >= 0, bot = h_high <= H - 1, left = w_low >= 0, right = w_high <= W - 1;
        if (top && left) load_vec8<__nv_bfloat16>(vl + ((long)h_low * W + w_low) * row_stride, v1);
        if (top && right) load_vec8<__nv_bfloat16>(vl + ((long)h_low * W + w_high) * row_stride, v2);
        if (bot && left) load_vec8<__nv_bfloat16>(vl + ((long)h_high * W + w_low) * row_stride, v3);
        if (bot && right) load_vec8<__nv_bfloat16>(vl + ((long)h_high * W + w_high) * row_stride, v4);
        const float w1 = hh * hw * aw, w2 = hh * lw * aw, w3 = lh * hw * aw, w4 = lh * lw * aw;
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] += w1 * v1[i] + w2 * v2[i] + w3 * v3[i] + w4 * v4[i];
      }
    }
  }
  uint4 o;
  o.x = pack_bf16(acc[0], acc[1]); o.y = pack_bf16(acc[2], acc[3]);
  o.z = pack_bf16(acc[4], acc[5]); o.w = pack_bf16(acc[6], acc[7]);
  *reinterpret_cast<uint4*>(p.out + item * p.D + part * 8) = o;
}


// ---- L = 4 levels x P = 4 points (the TESTR configuration), slimmed down.  ncu on the general kernel above
// (profiles/round1_summary.md): 4375 instructions per lane and item, 128 registers -> 16 resident warps per SM, issue
// slots 45 % busy, L2 6 % busy: bound by its own instruction stream, not by the gather.  Here: 16-byte loads of the
// offset / logit row, corner validity folded into the bilinear weights with clamped (always legal) addresses instead of
// per-corner branches and zero fills, 32-bit offsets inside an image, <= 64 registers for 4 CTAs per SM. ----
__device__ __forceinline__ void ldg256(const void* ptr, uint32_t (&r)[8]) {   // 32-byte aligned, read-only path
  asm("ld.global.nc.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
      : "l"(ptr));
}

template <typename PJ> struct ProjVec;
template <> struct ProjVec<float> {
  static constexpr int PER = 4;  // values per 16-byte load
  static __device__ __forceinline__ void load(const float* p, float (&f)[4]) {
    const float4 v = __ldg(reinterpret_cast<const float4*>(p));
    f[0] = v.x; f[1] = v.y; f[2] = v.z; f[3] = v.w;
  }
};
template <> struct ProjVec<__nv_bfloat16> {
  static constexpr int PER = 8;
  static __device__ __forceinline__ void load(const __nv_bfloat16* p, float (&f)[8]) {
    const uint4 q = __ldg(reinterpret_cast<const uint4*>(p));
    const float2 a = unpack_bf16(q.x), b = unpack_bf16(q.y), c = unpack_bf16(q.z), d = unpack_bf16(q.w);
    f[0] = a.x; f[1] = a.y; f[2] = b.x; f[3] = b.y; f[4] = c.x; f[5] = c.y; f[6] = d.x; f[7] = d.y;
  }
};

// CPL = channels per lane (8 or 16): with 16 the per-(query, head) work that does not depend on the channel — softmax,
// sampling coordinates, bilinear weights, addresses — is amortised over twice as many channels (2 lanes per item for D=32).
template <typename PJ, int CPL>
__global__ void __launch_bounds__(256, CPL == 8 ? 4 : 3) msda_fused44_kernel(const MsdaFusedParams p) {
  pdl_grid_sync();
  constexpr int LP = 16, PER = ProjVec<PJ>::PER, NV = CPL / 8;
  const long gtid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long item = gtid / p.lanes_per_item;
  const int part = (int)(gtid - item * p.lanes_per_item);
  if (item >= p.items) return;
  const int m = (int)(item % p.M);
  const long bq = item / p.M;
  const int b = (int)(bq / p.Lq);
  const int q = (int)(bq - (long)b * p.Lq);
  const PJ* row = reinterpret_cast<const PJ*>(p.proj) + bq * p.ldp;
  // softmax over the 16 logits of this (query, head)
  float w[LP];
#pragma unroll
  for (int i = 0; i < LP / PER; ++i) {
    float t[PER];
    ProjVec<PJ>::load(row + (long)p.M * LP * 2 + m * LP + i * PER, t);
#pragma unroll
    for (int j = 0; j < PER; ++j) w[i * PER + j] = t[j];
  }
  float mx = w[0];
#pragma unroll
  for (int i = 1; i < LP; ++i) mx = fmaxf(mx, w[i]);
  float sum = 0.f;
#pragma unroll
  for (int i = 0; i < LP; ++i) { w[i] = __expf(w[i] - mx); sum += w[i]; }
  const float inv = 1.f / sum;
  const float* refq = p.ref + (long)b * p.ref_batch_stride + (long)(q / p.q_per_ref) * 4 * p.ref_dim;
  const __nv_bfloat16* vbase = p.value + ((long)b * p.S * p.M + m) * p.D + part * CPL;
  const int row_stride = p.M * p.D;
  float acc[CPL];
#pragma unroll
  for (int i = 0; i < CPL; ++i) acc[i] = 0.f;
#pragma unroll
  for (int l = 0; l < 4; ++l) {
    const int H = (int)__ldg(p.shapes + 2 * l), W = (int)__ldg(p.shapes + 2 * l + 1);
    const __nv_bfloat16* vl = vbase + (long)__ldg(p.starts + l) * row_stride;
    const float fH = (float)H, fW = (float)W;
    const uint8_t* vlb = reinterpret_cast<const uint8_t*>(vl);   // byte offsets inside a level fit 32 bits
    const uint32_t rsb = (uint32_t)row_stride * 2u, wrs = (uint32_t)W * rsb;
    const float rx = __ldg(refq + l * p.ref_dim), ry = __ldg(refq + l * p.ref_dim + 1);
    float sx, sy;  // offset scale: 1/(W,H) for points, box_wh * 0.5 / P for boxes (ms_deform_attn.py:139-146)
    if (p.ref_dim == 2) { sx = 1.f / fW; sy = 1.f / fH; }
    else { sx = __ldg(refq + l * p.ref_dim + 2) * 0.125f; sy = __ldg(refq + l * p.ref_dim + 3) * 0.125f; }
    float off[8];  // (x, y) of the 4 points of this level
#pragma unroll
    for (int i = 0; i < 8 / PER; ++i) {
      float t[PER];
      ProjVec<PJ>::load(row + m * LP * 2 + l * 8 + i * PER, t);
#pragma unroll
      for (int j = 0; j < PER; ++j) off[i * PER + j] = t[j];
    }
#pragma unroll
    for (int s = 0; s < 4; ++s) {
      const float aw = w[l * 4 + s] * inv;
      const float h_im = (ry + off[2 * s + 1] * sy) * fH - 0.5f;
      const float w_im = (rx + off[2 * s] * sx) * fW - 0.5f;
      const bool valid = h_im > -1.f && w_im > -1.f && h_im < fH && w_im < fW;
      const float fh = floorf(h_im), fw = floorf(w_im);
      const int h_low = (int)fh, w_low = (int)fw;
      const float lh = h_im - fh, lw = w_im - fw;
      const float hh = 1.f - lh, hw = 1.f - lw;
      // a corner outside the map contributes zero: fold that into its weight and clamp its address into the map
      const float wt = (valid && h_low >= 0) ? hh * aw : 0.f;
      const float wb = (valid && h_low + 1 <= H - 1) ? lh * aw : 0.f;
      const float wl = (w_low >= 0) ? hw : 0.f;
      const float wr = (w_low + 1 <= W - 1) ? lw : 0.f;
      const int y0 = min(max(h_low, 0), H - 1), y1 = min(max(h_low + 1, 0), H - 1);
      const int x0 = min(max(w_low, 0), W - 1), x1 = min(max(w_low + 1, 0), W - 1);
      const uint32_t r0 = (uint32_t)y0 * wrs, r1 = (uint32_t)y1 * wrs, c0 = (uint32_t)x0 * rsb, c1 = (uint32_t)x1 * rsb;
      const float w1 = wt * wl, w2 = wt * wr, w3 = wb * wl, w4 = wb * wr;
      if constexpr (CPL == 16) {
        if (p.wide_loads) {
          // one 32-byte load per corner (LDG.256, sm_100): this lane's 16 channels are exactly one sector, so the gather
          // issues half the L1 requests of two 16-byte loads (each of which touched the sector for half of its bytes)
          uint32_t a1[8], a2[8], a3[8], a4[8];
          ldg256(vlb + (r0 + c0), a1);
          ldg256(vlb + (r0 + c1), a2);
          ldg256(vlb + (r1 + c0), a3);
          ldg256(vlb + (r1 + c1), a4);
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            const float2 v1 = unpack_bf16(a1[i]), v2 = unpack_bf16(a2[i]), v3 = unpack_bf16(a3[i]), v4 = unpack_bf16(a4[i]);
            acc[2 * i] = fmaf(w4, v4.x, fmaf(w3, v3.x, fmaf(w2, v2.x, fmaf(w1, v1.x, acc[2 * i]))));
            acc[2 * i + 1] = fmaf(w4, v4.y, fmaf(w3, v3.y, fmaf(w2, v2.y, fmaf(w1, v1.y, acc[2 * i + 1]))));
          }
          continue;
        }
      }
#pragma unroll
      for (int nv = 0; nv < NV; ++nv) {
        const uint4 q1 = __ldg(reinterpret_cast<const uint4*>(vlb + (r0 + c0)) + nv);
        const uint4 q2 = __ldg(reinterpret_cast<const uint4*>(vlb + (r0 + c1)) + nv);
        const uint4 q3 = __ldg(reinterpret_cast<const uint4*>(vlb + (r1 + c0)) + nv);
        const uint4 q4 = __ldg(reinterpret_cast<const uint4*>(vlb + (r1 + c1)) + nv);
        const uint32_t a1[4] = {q1.x, q1.y, q1.z, q1.w}, a2[4] = {q2.x, q2.y, q2.z, q2.w};
        const uint32_t a3[4] = {q3.x, q3.y, q3.z, q3.w}, a4[4] = {q4.x, q4.y, q4.z, q4.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float2 v1 = unpack_bf16(a1[i]), v2 = unpack_bf16(a2[i]), v3 = unpack_bf16(a3[i]), v4 = unpack_bf16(a4[i]);
          const int c = nv * 8 + 2 * i;
          acc[c] = fmaf(w4, v4.x, fmaf(w3, v3.x, fmaf(w2, v2.x, fmaf(w1, v1.x, acc[c]))));
          acc[c + 1] = fmaf(w4, v4.y, fmaf(w3, v3.y, fmaf(w2, v2.y, fmaf(w1, v1.y, acc[c + 1]))));
        }
      }
    }
  }
#pragma unroll
  for (int nv = 0; nv < NV; ++nv) {
    uint4 o;
    o.x = pack_bf16(acc[nv * 8 + 0], acc[nv * 8 + 1]); o.y = pack_bf16(acc[nv * 8 + 2], acc[nv * 8 + 3]);
    o.z = pack_bf16(acc[nv * 8 + 4], acc[nv * 8 + 5]); o.w = pack_bf16(acc[nv * 8 + 6], acc[nv * 8 + 7]);
    *reinterpret_cast<uint4*>(p.out + item * p.D + part * CPL + nv * 8) = o;
  }
}

}  // namespace
}  // namespace tair

using namespace tair;

extern "C" int tair_msda_forward(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                                 const float* sampling_loc, const float* attn_weight, void* out, int32_t B,
                                 int32_t S, int32_t M, int32_t D, int32_t L, int32_t Lq, int32_t P,
                                 int32_t value_bf16, int32_t out_bf16, void* stream) {
  TAIR_REQUIRE(value && spatial_shapes && level_start_index && sampling_loc && attn_weight && out,
               "msda_forward: NULL pointer");
  TAIR_REQUIRE(B > 0 && S > 0 && M > 0 && D > 0 && L > 0 && Lq > 0 && P > 0, "msda_forward: bad shape");
  TAIR_REQUIRE(D % 8 == 0 && D <= 256 && (32 % (D / 8)) == 0,
               "msda_forward: channels per head must be 8,16,32,64,128 or 256 (got %d)", D);
  TAIR_REQUIRE((reinterpret_cast<uintptr_t>(value) % 32) == 0 && (reinterpret_cast<uintptr_t>(out) % 32) == 0 &&
                   (reinterpret_cast<uintptr_t>(sampling_loc) % 8) == 0,
               "msda_forward: tensors must be 32-byte aligned");
  MsdaParams p{};
  p.value = value; p.shapes = spatial_shapes; p.starts = level_start_index;
  p.loc = sampling_loc; p.attw = attn_weight; p.out = out;
  p.B = B; p.S = S; p.M = M; p.D = D; p.L = L; p.Lq = Lq; p.P = P;
  p.lanes_per_item = D / 8;
  p.items = (long)B * Lq * M;
  const long threads_total = p.items * p.lanes_per_item;
  const int block = 256;
  const long grid = (threads_total + block - 1) / block;
  TAIR_REQUIRE(grid < (1l << 31), "msda_forward: problem too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (value_bf16 && out_bf16) TAIR_LAUNCH((msda_fwd_kernel<__nv_bfloat16, __nv_bfloat16>), (unsigned)grid, block, 0, st, p);
  else if (value_bf16) TAIR_LAUNCH((msda_fwd_kernel<__nv_bfloat16, float>), (unsigned)grid, block, 0, st, p);
  else if (out_bf16) TAIR_LAUNCH((msda_fwd_kernel<float, __nv_bfloat16>), (unsigned)grid, block, 0, st, p);
  else TAIR_LAUNCH((msda_fwd_kernel<float, float>), (unsigned)grid, block, 0, st, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("msda_fwd_kernel");
}

extern "C" int tair_msda_fused(const void* value, const int64_t* spatial_shapes, const int64_t* level_start_index,
                               const void* proj, int64_t ldp, int32_t proj_bf16, const float* ref, int32_t ref_dim,
                               int64_t ref_batch_stride, int32_t q_per_ref, void* out, int32_t B, int32_t S,
                               int32_t M, int32_t D, int32_t L, int32_t Lq, int32_t P, void* stream) {
  TAIR_REQUIRE(value && spatial_shapes && level_start_index && proj && ref && out, "msda_fused: NULL pointer");
  TAIR_REQUIRE(B > 0 && S > 0 && M > 0 && D > 0 && L > 0 && Lq > 0 && P > 0 && q_per_ref > 0, "msda_fused: bad shape");
  TAIR_REQUIRE(ref_dim == 2 || ref_dim == 4, "msda_fused: reference points must have 2 or 4 coordinates");
  TAIR_REQUIRE(L * P <= MSDA_MAX_LP && L <= 8 && P <= 8, "msda_fused: needs L <= 8, P <= 8, L*P <= %d", MSDA_MAX_LP);
  TAIR_REQUIRE(D % 8 == 0 && D <= 256 && (32 % (D / 8)) == 0, "msda_fused: unsupported channels per head %d", D);
  TAIR_REQUIRE(ldp >= (int64_t)M * L * P * 3 && ldp % 2 == 0, "msda_fused: projection row too short");
  TAIR_REQUIRE((reinterpret_cast<uintptr_t>(value) % 16) == 0 && (reinterpret_cast<uintptr_t>(out) % 16) == 0 &&
                   (reinterpret_cast<uintptr_t>(proj) % 4) == 0, "msda_fused: misaligned tensor");
  MsdaFusedParams p{};
  p.value = reinterpret_cast<const __nv_bfloat16*>(value);
  p.shapes = spatial_shapes; p.starts = level_start_index;
  p.proj = proj; p.ldp = ldp; p.ref = ref; p.ref_dim = ref_dim; p.ref_batch_stride = ref_batch_stride;
  p.q_per_ref = q_per_ref; p.out = reinterpret_cast<__nv_bfloat16*>(out);
  p.B = B; p.S = S; p.M = M; p.D = D; p.L = L; p.Lq = Lq; p.P = P;
  p.lanes_per_item = D / 8;
  p.items = (long)B * Lq * M;
  // 32-byte loads need 32-byte aligned taps: value base and the per-head row (D * 2 bytes) multiples of 32
  p.wide_loads = ((reinterpret_cast<uintptr_t>(value) % 32) == 0 && (D * 2) % 32 == 0 && !getenv("TAIR_MSDA_LDG128")) ? 1 : 0;
  const long threads_total = p.items * p.lanes_per_item;
  const long grid = (threads_total + 255) / 256;
  TAIR_REQUIRE(grid < (1l << 31), "msda_fused: problem too large");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const size_t pj = proj_bf16 ? 2 : 4;
  const bool fast44 = L == 4 && P == 4 && (reinterpret_cast<uintptr_t>(proj) % 16) == 0 && (ldp * pj) % 16 == 0 &&
                      ((size_t)M * 32 * pj) % 16 == 0 && (long)S * M * D < (1l << 30) && !getenv("TAIR_MSDA_GENERIC");
  if (fast44 && D % 16 == 0 && !getenv("TAIR_MSDA_CPL8")) {
    p.lanes_per_item = D / 16;
    const long grid16 = (p.items * p.lanes_per_item + 255) / 256;
    if (proj_bf16) TAIR_LAUNCH((msda_fused44_kernel<__nv_bfloat16, 16>), (unsigned)grid16, 256, 0, st, p);
    else TAIR_LAUNCH((msda_fused44_kernel<float, 16>), (unsigned)grid16, 256, 0, st, p);
  } else if (fast44) {
    if (proj_bf16) TAIR_LAUNCH((msda_fused44_kernel<__nv_bfloat16, 8>), (unsigned)grid, 256, 0, st, p);
    else TAIR_LAUNCH((msda_fused44_kernel<float, 8>), (unsigned)grid, 256, 0, st, p);
  } else if (L == 4 && P == 4) {
    if (proj_bf16) TAIR_LAUNCH((msda_fused_kernel<4, 4, __nv_bfloat16>), (unsigned)grid, 256, 0, st, p);
    else TAIR_LAUNCH((msda_fused_kernel<4, 4, float>), (unsigned)grid, 256, 0, st, p);
  } else {
    if (proj_bf16) TAIR_LAUNCH((msda_fused_kernel<0, 0, __nv_bfloat16>), (unsigned)grid, 256, 0, st, p);
    else TAIR_LAUNCH((msda_fused_kernel<0, 0, float>), (unsigned)grid, 256, 0, st, p);
  }
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("msda_fused_kernel");
}

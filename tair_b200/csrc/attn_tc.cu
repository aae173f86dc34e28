// Flash-style attention for sm_100a, head_dim 64, no mask (UNet self/cross attention).
//
// Replaces F.scaled_dot_product_attention at terediff/model/attention.py:206
// (SDPCrossAttention.forward :189-216): softmax(Q K^T * scale) V per (batch, head).
//
// One CTA = one (batch, head, 128-query tile); two CTAs are co-resident per SM.
//   warp 0       TMA producer: Q once, then K blocks (128 keys) into a 3-stage ring and V blocks into a 2-stage ring
//   warp 1       TMEM owner + MMA issuer: S = Q K^T (SS, M128 N128 K64) and O += P V (TS: P read from TMEM)
//   warps 2..5   softmax: one thread per query row; S read with tcgen05.ld; P = exp2(S*c - m) written as packed bf16
//                with tcgen05.st into its own TMEM columns
// Software pipeline: as soon as the softmax threads hold S(j) in registers (`s_free`) the MMA warp issues
// S(j+1) = Q K(j+1)^T, which therefore runs UNDER the exponentials of block j; P(j) V follows when P(j) is written.
// The softmax warps then go from block to block without waiting for the tensor pipe (the earlier layout aliased P
// over S, which serialised softmax(j) -> P V(j) -> S(j+1) and left the MUFU pipe 40 % idle, profiles/round1_summary.md).
// O accumulates in TMEM across key blocks; the running maximum is updated lazily (only when a block's maximum
// exceeds it by more than 2^8, so exp2 arguments stay <= 8), in which case the softmax threads rescale O in TMEM.
// TMEM (256 columns): S fp32 [0,128) | P bf16x2 [128,192) | O fp32 [192,256).
#include <atomic>
#include <math_constants.h>

#include "../../include/tair_b200.h"
#include <cstdlib>

#include "common.cuh"

namespace tair {
extern std::atomic<int64_t> g_launch_count;
namespace {

constexpr int AT_BM = 128;   // queries per CTA
constexpr int AT_BN = 128;   // keys per block
constexpr int AT_D = 64;     // head dim
constexpr int AT_THREADS = 192;
constexpr uint32_t TILE_BYTES = 128 * 64 * 2;  // 16 KB: one 128-row x 64-col bf16 tile
constexpr uint32_t SM_Q = 0;
constexpr int K_STAGES = 3, V_STAGES = 2;
constexpr uint32_t SM_K = SM_Q + TILE_BYTES;
constexpr uint32_t SM_V = SM_K + K_STAGES * TILE_BYTES;
constexpr uint32_t SM_BAR = SM_V + V_STAGES * TILE_BYTES;
constexpr uint32_t AT_SMEM = SM_BAR + 144;
constexpr uint32_t AT_TMEM_COLS = 256;

struct AttnParams {
  int H, Lq, Lk;
  int n_inner;       // sequences are indexed (outer, inner); plain batched attention has n_inner == 1
  int causal;        // key j attends only to queries i >= j (OpenCLIP text transformer)
  int group;         // > 0: block-diagonal mode — the rows are back-to-back sequences of `group` tokens, a tile packs
                     // tile_rows / group of them and every row only sees the keys of its own sequence
  int tile_rows;     // query rows per CTA: 128, or (128 / group) * group in block-diagonal mode
  const float* bias; // block-diagonal mode only: additive table [bias_nw][H][group (key)][group (query)] in units of
  int bias_nw;       // 1/scale (S + table, then * scale); sequence w uses table w % bias_nw (Swin window masks)
  int q_tiles, nblk;
  float scale_log2;  // scale * log2(e)
  __nv_bfloat16* o;
  int64_t ldo;
  int64_t o_outer, o_inner, o_tok;  // output row = outer*o_outer + inner*o_inner + token*o_tok
};


// P = exp2(S c - m) for one 128-wide S row held in registers -> packed bf16 pairs + four partial row sums.  Two
// straight-line instantiations: the unmasked one (every block but the last / the causal diagonal / packed groups) carries
// no per-element compare + select.  (Measured and rejected in round 2, tools/attn_probe.py + tools/probes/mufu_probe.cu:
// forcing a FFMA / MUFU interleave with volatile asm — ptxas reschedules anyway, no change; packing to bf16 with integer
// rounding + PRMT instead of F2FP — 473 vs 456 us; two query tiles per CTA whose softmax warps hand the MUFU to each other
// through named barriers — 482 vs 456 us: one warp alone sustains one exponential per 9.4 cycles, two overlapping 8.3.)
template <bool MASKED>
__device__ __forceinline__ void softmax_exp_row(uint32_t (&sv)[AT_BN], float c, float m_run, int klo, int kvalid,
                                                uint32_t (&pk)[AT_BN / 2], float& ls0, float& ls1, float& ls2, float& ls3) {
#pragma unroll
  for (int i = 0; i < AT_BN; i += 4) {
    float e0 = ex2_approx(fmaf(__uint_as_float(sv[i + 0]), c, -m_run));
    float e1 = ex2_approx(fmaf(__uint_as_float(sv[i + 1]), c, -m_run));
    float e2 = ex2_approx(fmaf(__uint_as_float(sv[i + 2]), c, -m_run));
    float e3 = ex2_approx(fmaf(__uint_as_float(sv[i + 3]), c, -m_run));
    if (MASKED) {
      if (i + 0 >= kvalid || i + 0 < klo) e0 = 0.f;
      if (i + 1 >= kvalid || i + 1 < klo) e1 = 0.f;
      if (i + 2 >= kvalid || i + 2 < klo) e2 = 0.f;
      if (i + 3 >= kvalid || i + 3 < klo) e3 = 0.f;
    }
    ls0 += e0; ls1 += e1; ls2 += e2; ls3 += e3;
    pk[(i >> 1) + 0] = pack_bf16(e0, e1);
    pk[(i >> 1) + 1] = pack_bf16(e2, e3);
  }
}

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_tc_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
               const __grid_constant__ CUtensorMap tmV, const AttnParams p) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + SM_BAR;
  const uint32_t q_full = bar + 0;
  const uint32_t s_full = bar + 8;    // MMA -> softmax: S(j) complete
  const uint32_t s_free = bar + 16;   // softmax -> MMA: S(j) is in registers, the S columns may be overwritten
  const uint32_t p_full = bar + 24;   // softmax -> MMA: P(j) written
  const uint32_t p_free = bar + 32;   // MMA -> softmax: P V(j) retired (P columns and O are free again)
  const uint32_t o_done = bar + 40;
  auto k_full = [&](int s) { return bar + 48 + 8u * s; };
  auto k_empty = [&](int s) { return bar + 72 + 8u * s; };
  auto v_full = [&](int s) { return bar + 96 + 8u * s; };
  auto v_empty = [&](int s) { return bar + 112 + 8u * s; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + SM_BAR + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int qt = blockIdx.x % p.q_tiles;
  const int bh = blockIdx.x / p.q_tiles;
  const int h = bh % p.H;
  const int seq = bh / p.H;
  const int s_in = seq % p.n_inner, s_out = seq / p.n_inner;
  const int q0 = qt * p.tile_rows;
  const int kv0 = p.group > 0 ? q0 : 0;   // block-diagonal mode: the only key block is the tile's own rows

  if (threadIdx.x == 0) {
    if ((sbase & 1023u) != 0) __trap();  // swizzle-128B tiles need 1024-byte alignment
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(p_full, 4);
    mbar_init(p_free, 1);
    for (int s = 0; s < K_STAGES; ++s) {
      mbar_init(k_full(s), 1);
      mbar_init(k_empty(s), 1);
    }
    for (int s = 0; s < V_STAGES; ++s) {
      mbar_init(v_full(s), 1);
      mbar_init(v_empty(s), 1);
    }
    mbar_init(o_done, 1);
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_grid_sync();   // barrier / tensor-memory set-up above overlaps the previous kernel's tail
  const uint32_t tm_S = tmem;
  const uint32_t tm_P = tmem + 128;
  const uint32_t tm_O = tmem + 192;

  // Warps 0 and 1 run warp-uniform loops and ONE elected lane issues the TMA / tcgen05 instructions, so that ptxas keeps
  // their operands in uniform registers (a `lane == 0` branch turns each issue into a vote/elect/R2UR loop).
  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(q_full, TILE_BYTES);
      tma_load_4d(sbase + SM_Q, &tmQ, q_full, h * AT_D, q0, s_in, s_out);
    }
    __syncwarp();
    for (int j = 0; j < p.nblk; ++j) {
      const int ks = j % K_STAGES, vs = j % V_STAGES;
      mbar_wait(k_empty(ks), ((j / K_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(k_full(ks), TILE_BYTES);
        tma_load_4d(sbase + SM_K + ks * TILE_BYTES, &tmK, k_full(ks), h * AT_D, kv0 + j * AT_BN, s_in, s_out);
      }
      __syncwarp();
      mbar_wait(v_empty(vs), ((j / V_STAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(v_full(vs), TILE_BYTES);
        tma_load_4d(sbase + SM_V + vs * TILE_BYTES, &tmV, v_full(vs), h * AT_D, kv0 + j * AT_BN, s_in, s_out);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, 128, 0, 0);  // S: A=Q K-major, B=K K-major
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);   // O: A=P K-major, B=V MN-major
    constexpr uint32_t desc_hi = umma_desc_hi_sw128(1024);
    const uint32_t q_lo = umma_desc_lo(sbase + SM_Q, 16);
    auto issue_s = [&](int j) {   // S(j) = Q K(j)^T; releases the K stage and signals s_full when it retires
      const int ks = j % K_STAGES;
      mbar_wait(k_full(ks), (j / K_STAGES) & 1);
      tc_fence_after();
      const uint32_t k_lo = umma_desc_lo(sbase + SM_K + ks * TILE_BYTES, 16);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k)
          umma_ss_lohi(tm_S, q_lo + 2 * k, k_lo + 2 * k, desc_hi, idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(k_empty(ks));
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    issue_s(0);
    for (int j = 0; j < p.nblk; ++j) {
      if (j + 1 < p.nblk) {
        mbar_wait(s_free, j & 1);   // the softmax threads hold S(j) in registers
        tc_fence_after();
        issue_s(j + 1);             // runs while they compute the exponentials of block j
      }
      const int vs = j % V_STAGES;
      mbar_wait(p_full, j & 1);
      mbar_wait(v_full(vs), (j / V_STAGES) & 1);
      tc_fence_after();
      const uint32_t v_lo = umma_desc_lo(sbase + SM_V + vs * TILE_BYTES, 16);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_BN / 16; ++k)  // A = P from TMEM: 16 bf16 of K per step = 8 columns; V advances 16 rows
          umma_ts_lohi(tm_O, tm_P + k * 8, v_lo + k * (2048 >> 4), desc_hi, idesc_o, (j | k) != 0);
        umma_commit(v_empty(vs));
        umma_commit(p_free);
        if (j + 1 == p.nblk) umma_commit(o_done);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;            // TMEM lane quadrant this warp may access
    const int r = quad * 32 + lane;       // query row inside the tile
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    float m_run = -CUDART_INF_F, l_run = 0.f;
    const float c = p.scale_log2;

    for (int j = 0; j < p.nblk; ++j) {
      mbar_wait(s_full, j & 1);   // S(j) is complete
      tc_fence_after();
      int klo = 0;                           // this row sees key columns [klo, kvalid) of the block
      int kvalid = p.Lk - kv0 - j * AT_BN;   // columns >= kvalid are padding (only in the last block)
      if (p.causal) {                        // ... or lie above the diagonal for this query row
        const int lim = q0 + r + 1 - j * AT_BN;
        kvalid = lim < kvalid ? lim : kvalid;
      }
      if (p.group > 0) {                     // ... or belong to another sequence of the packed tile
        klo = (r / p.group) * p.group;
        const int hi = klo + p.group;
        if (r >= p.tile_rows || klo >= kvalid) { klo = 0; kvalid = 1; }   // unused row: keep the arithmetic finite
        else kvalid = hi < kvalid ? hi : kvalid;
      }
      // warp-uniform: the tcgen05.ld/st below are .sync.aligned and must not sit in a divergent branch
      const bool full = __all_sync(0xffffffffu, klo == 0 && kvalid >= AT_BN);
      // the whole S row (128 fp32) is pulled into registers with four back-to-back tcgen05.ld and ONE wait, and is
      // used for both the max and the exponentials (the first version re-read TMEM and stalled on 8 waits per block)
      uint32_t sv[AT_BN];
#pragma unroll
      for (int cc = 0; cc < AT_BN; cc += 32) {
        uint32_t (&chunk)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[cc]);
        tmem_ld_32x32(tm_S + lane_off + cc, chunk);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);   // S(j) now lives in registers: the MMA warp may start S(j+1)
      if (p.bias != nullptr) {     // relative-position bias (+ shifted-window mask): lanes = consecutive query rows
        const int gsz = p.group;
        const int wq = (q0 + r) / gsz;                       // global sequence (window) index of this row
        const float* tb = p.bias + (((int64_t)(wq % p.bias_nw) * p.H + h) * gsz) * gsz + (r - klo);
#pragma unroll
        for (int i = 0; i < AT_BN; ++i)
          if (i >= klo && i < kvalid) sv[i] = __float_as_uint(__uint_as_float(sv[i]) + __ldg(tb + (i - klo) * gsz));
      }
      float mx;
      {
        float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
        if (full) {
#pragma unroll
          for (int i = 0; i < AT_BN; i += 4) {
            m0 = fmaxf(m0, __uint_as_float(sv[i]));
            m1 = fmaxf(m1, __uint_as_float(sv[i + 1]));
            m2 = fmaxf(m2, __uint_as_float(sv[i + 2]));
            m3 = fmaxf(m3, __uint_as_float(sv[i + 3]));
          }
        } else {
#pragma unroll
          for (int i = 0; i < AT_BN; ++i)
            if (i >= klo && i < kvalid) m0 = fmaxf(m0, __uint_as_float(sv[i]));
        }
        mx = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3));
      }
      // lazy running maximum: move it only when this block exceeds it by more than 8 (in log2 units)
      const float m_blk = mx * c;
      const bool bump = m_blk > m_run + 8.f;
      const bool any_bump = __any_sync(0xffffffffu, bump);
      float alpha = 1.f;
      if (any_bump) {
        const float m_new = bump ? m_blk : m_run;
        alpha = ex2_approx(m_run - m_new);   // 1 for lanes that keep their maximum, 0 on the first block
        l_run *= alpha;
        m_run = m_new;
      }
      // P = exp2(S*c - m) -> packed bf16, kept in REGISTERS (the S values die as they are consumed) and written to the P
      // columns only after P V(j-1) has retired: the tensor pipe is shared with the co-resident CTA, so that MMA can sit
      // behind 400+ cycles of foreign work, and waiting for it before the exponentials left the MUFU idle.
      // Two straight-line copies: the unmasked one (all blocks but the last / the causal diagonal) carries no
      // per-element compare+select — they were 2 of 7 instructions per element when the mask was predicated in.
      float ls0 = 0.f, ls1 = 0.f, ls2 = 0.f, ls3 = 0.f;
      uint32_t pk[AT_BN / 2];
      if (full) softmax_exp_row<false>(sv, c, m_run, klo, kvalid, pk, ls0, ls1, ls2, ls3);
      else softmax_exp_row<true>(sv, c, m_run, klo, kvalid, pk, ls0, ls1, ls2, ls3);
      if (j > 0) {                // P V(j-1) must have retired before O is rescaled or P is overwritten
        mbar_wait(p_free, (j - 1) & 1);
        tc_fence_after();
        if (any_bump) {           // rescale the accumulated O row in TMEM
#pragma unroll
          for (int half = 0; half < 4; ++half) {
            uint32_t ov[16];
            tmem_ld_32x16(tm_O + lane_off + half * 16, ov);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i) ov[i] = __float_as_uint(__uint_as_float(ov[i]) * alpha);
            tmem_st_32x16(tm_O + lane_off + half * 16, ov);
          }
        }
      }
#pragma unroll
      for (int cc = 0; cc < AT_BN / 2; cc += 16) {
        uint32_t (&chunk)[16] = *reinterpret_cast<uint32_t (*)[16]>(&pk[cc]);
        tmem_st_32x16(tm_P + lane_off + cc, chunk);
      }
      l_run += (ls0 + ls1) + (ls2 + ls3);
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full);
    }
    mbar_wait(o_done, 0);
    tc_fence_after();
    const int q = q0 + r;
    const float inv = 1.f / l_run;
    __nv_bfloat16* op = p.o + ((int64_t)s_out * p.o_outer + (int64_t)s_in * p.o_inner + (int64_t)q * p.o_tok) * p.ldo + h * AT_D;
#pragma unroll
    for (int half = 0; half < 2; ++half) {
      uint32_t ov[32];
      tmem_ld_32x32(tm_O + lane_off + half * 32, ov);
      tmem_ld_wait();
      if (q < p.Lq && r < p.tile_rows) {
#pragma unroll
        for (int i = 0; i < 32; i += 8) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(ov[i + 0]) * inv, __uint_as_float(ov[i + 1]) * inv);
          w.y = pack_bf16(__uint_as_float(ov[i + 2]) * inv, __uint_as_float(ov[i + 3]) * inv);
          w.z = pack_bf16(__uint_as_float(ov[i + 4]) * inv, __uint_as_float(ov[i + 5]) * inv);
          w.w = pack_bf16(__uint_as_float(ov[i + 6]) * inv, __uint_as_float(ov[i + 7]) * inv);
          *reinterpret_cast<uint4*>(op + half * 32 + i) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, AT_TMEM_COLS);
  }
}


// ---------------------------------------------------------------------------------------------------------------------
// Key/value-stationary variant for SHORT contexts (the UNet / ControlNet cross-attention: Lk = 77 CLIP tokens,
// attention.py:189-216 with context).  With one key block per query tile the flash kernel above degenerates into
// 2560 CTAs that each pay barrier set-up, tensor-memory allocation and three dependent TMA round trips for ~1 us of
// math (67 us per launch at 4096 x 77 against a 13 us HBM floor).  Here a CTA keeps K and V of one (batch, head) resident
// and STREAMS query tiles through a 3-deep ring:
//   warp 0      TMA: K, V once; Q tiles into the ring
//   warp 1      MMA: S(i) = Q(i) K^T (M128 N80 K64) and O(i) = P(i) V (TS form, K = 80 keys)
//   warps 2..5  softmax of tile i+1, THEN the drain of O(i): P is double buffered so that P(i+1) is written while
//               P(i) V is still running, and the tensor pipe computes O(i+1) under the softmax of tile i+2
// TMEM (256 columns): S fp32 [0,80) | P bf16x2 [80,120) and [120,160) | O fp32 [160,224).  A single key block needs no
// online rescaling: P = exp2(S c - max c), O / sum(P).
constexpr int KS_NK = 80;          // key columns per MMA (77 valid, zero rows from TMA's out-of-bounds fill)
constexpr int KS_QSTAGES = 3;
constexpr uint32_t KS_SM_K = 0;
constexpr uint32_t KS_SM_V = KS_SM_K + TILE_BYTES;
constexpr uint32_t KS_SM_Q = KS_SM_V + TILE_BYTES;
constexpr uint32_t KS_SM_STG = KS_SM_Q + KS_QSTAGES * TILE_BYTES;   // 4 warps x (32 rows x 128 B) output staging
constexpr uint32_t KS_SM_BAR = KS_SM_STG + 4 * 4096;
constexpr uint32_t KS_SMEM = KS_SM_BAR + 160;

__global__ void __launch_bounds__(AT_THREADS, 2)
attn_kvs_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                const __grid_constant__ CUtensorMap tmV, const __grid_constant__ CUtensorMap tmO, const AttnParams p,
                int tiles_per_cta, int chunks) {
  extern __shared__ __align__(1024) uint8_t smem[];
  const uint32_t sbase = smem_u32(smem);
  const uint32_t bar = sbase + KS_SM_BAR;
  const uint32_t kv_full = bar + 0, s_full = bar + 8, s_free = bar + 16, o_full = bar + 24, o_free = bar + 32;
  auto p_full = [&](int b) { return bar + 40 + 8u * b; };
  auto q_full = [&](int s) { return bar + 56 + 8u * s; };
  auto q_empty = [&](int s) { return bar + 80 + 8u * s; };
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + KS_SM_BAR + 128);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int chunk = blockIdx.x % chunks;
  const int bh = blockIdx.x / chunks;
  const int h = bh % p.H, seq = bh / p.H;
  const int t0 = chunk * tiles_per_cta;
  int nt = p.q_tiles - t0;
  nt = nt < tiles_per_cta ? nt : tiles_per_cta;   // >= 1 by construction of the grid

  if (threadIdx.x == 0) {
    if ((sbase & 1023u) != 0) __trap();
    mbar_init(kv_full, 1);
    mbar_init(s_full, 1);
    mbar_init(s_free, 4);
    mbar_init(o_full, 1);
    mbar_init(o_free, 4);
    mbar_init(p_full(0), 4);
    mbar_init(p_full(1), 4);
    for (int s = 0; s < KS_QSTAGES; ++s) {
      mbar_init(q_full(s), 1);
      mbar_init(q_empty(s), 1);
    }
    mbar_fence_init();
  }
  if (warp == 1) {
    tmem_alloc(smem_u32(tmem_slot), AT_TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_grid_sync();
  const uint32_t tm_S = tmem, tm_P = tmem + 80, tm_O = tmem + 160;

  if (warp == 0) {
    if (elect_one()) {
      tma_prefetch_desc(&tmQ);
      tma_prefetch_desc(&tmK);
      tma_prefetch_desc(&tmV);
      mbar_expect_tx(kv_full, 2 * TILE_BYTES);
      tma_load_4d(sbase + KS_SM_K, &tmK, kv_full, h * AT_D, 0, 0, seq);
      tma_load_4d(sbase + KS_SM_V, &tmV, kv_full, h * AT_D, 0, 0, seq);
    }
    __syncwarp();
    for (int i = 0; i < nt; ++i) {
      const int qs = i % KS_QSTAGES;
      mbar_wait(q_empty(qs), ((i / KS_QSTAGES) & 1) ^ 1);
      if (elect_one()) {
        mbar_expect_tx(q_full(qs), TILE_BYTES);
        tma_load_4d(sbase + KS_SM_Q + qs * TILE_BYTES, &tmQ, q_full(qs), h * AT_D, (t0 + i) * AT_BM, 0, seq);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = umma_idesc_bf16(128, KS_NK, 0, 0);
    constexpr uint32_t idesc_o = umma_idesc_bf16(128, 64, 0, 1);
    constexpr uint32_t desc_hi = umma_desc_hi_sw128(1024);
    const uint32_t k_lo = umma_desc_lo(sbase + KS_SM_K, 16);
    const uint32_t v_lo = umma_desc_lo(sbase + KS_SM_V, 16);
    auto issue_s = [&](int i) {
      const int qs = i % KS_QSTAGES;
      mbar_wait(q_full(qs), (i / KS_QSTAGES) & 1);
      tc_fence_after();
      const uint32_t q_lo = umma_desc_lo(sbase + KS_SM_Q + qs * TILE_BYTES, 16);
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < AT_D / 16; ++k) umma_ss_lohi(tm_S, q_lo + 2 * k, k_lo + 2 * k, desc_hi, idesc_s, k != 0);
        umma_commit(s_full);
        umma_commit(q_empty(qs));
      }
      __syncwarp();
    };
    mbar_wait(kv_full, 0);
    issue_s(0);
    for (int i = 0; i < nt; ++i) {
      if (i + 1 < nt) {
        mbar_wait(s_free, i & 1);     // S(i) is in the softmax threads' registers
        tc_fence_after();
        issue_s(i + 1);
      }
      mbar_wait(p_full(i & 1), (i >> 1) & 1);
      if (i > 0) mbar_wait(o_free, (i - 1) & 1);   // O(i-1) has been drained
      tc_fence_after();
      if (elect_one()) {
#pragma unroll
        for (int k = 0; k < KS_NK / 16; ++k)
          umma_ts_lohi(tm_O, tm_P + (i & 1) * 40 + k * 8, v_lo + k * (2048 >> 4), desc_hi, idesc_o, k != 0);
        umma_commit(o_full);
      }
      __syncwarp();
    }
  } else {
    const int quad = warp & 3;
    const int r = quad * 32 + lane;
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const float c = p.scale_log2;
    const int kvalid = p.Lk;            // <= KS_NK
    float inv_l[2] = {0.f, 0.f};        // 1 / row sum of tiles i (slot i & 1)

    // O(i) / l(i) -> bf16 -> 128B-swizzled staging rows -> ONE bulk tensor store per warp (32 rows x 64 columns; rows past
    // Lq are clipped by the hardware).  Thread-per-row 16-byte global stores touch 32 different lines per instruction: with
    // a drain per tile they kept the LSU busier than the exponentials keep the MUFU.
    uint8_t* stg = smem + KS_SM_STG + quad * 4096;
    const uint32_t stg_u32 = smem_u32(stg);
    const int sw7 = lane & 7;
    auto drain = [&](int i) {
      mbar_wait(o_full, i & 1);
      tc_fence_after();
      const float inv = inv_l[i & 1];
      uint32_t ov[2][32];
      tmem_ld_32x32(tm_O + lane_off, ov[0]);
      tmem_ld_32x32(tm_O + lane_off + 32, ov[1]);
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(o_free);   // O is in registers: the next P V may overwrite the accumulator
      if (elect_one()) tma_store_wait_read<0>();   // the previous store of this warp has left the staging buffer
      __syncwarp();
#pragma unroll
      for (int half = 0; half < 2; ++half)
#pragma unroll
        for (int j = 0; j < 32; j += 8) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(ov[half][j + 0]) * inv, __uint_as_float(ov[half][j + 1]) * inv);
          w.y = pack_bf16(__uint_as_float(ov[half][j + 2]) * inv, __uint_as_float(ov[half][j + 3]) * inv);
          w.z = pack_bf16(__uint_as_float(ov[half][j + 4]) * inv, __uint_as_float(ov[half][j + 5]) * inv);
          w.w = pack_bf16(__uint_as_float(ov[half][j + 6]) * inv, __uint_as_float(ov[half][j + 7]) * inv);
          *reinterpret_cast<uint4*>(stg + lane * 128 + ((((half * 32 + j) >> 3) ^ sw7) << 4)) = w;
        }
      fence_async_smem();
      __syncwarp();
      if (elect_one()) {
        asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];\n" ::"l"(&tmO),
                     "r"(stg_u32), "r"(h * AT_D), "r"((t0 + i) * AT_BM + quad * 32), "r"(0), "r"(seq)
                     : "memory");
        tma_store_commit();
      }
    };

    for (int i = 0; i < nt; ++i) {
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      uint32_t sv[KS_NK];
      {
        uint32_t (&c0)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[0]);
        uint32_t (&c1)[32] = *reinterpret_cast<uint32_t (*)[32]>(&sv[32]);
        uint32_t (&c2)[16] = *reinterpret_cast<uint32_t (*)[16]>(&sv[64]);
        tmem_ld_32x32(tm_S + lane_off, c0);
        tmem_ld_32x32(tm_S + lane_off + 32, c1);
        tmem_ld_32x16(tm_S + lane_off + 64, c2);
      }
      tmem_ld_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(s_free);
      float m0 = -CUDART_INF_F, m1 = -CUDART_INF_F, m2 = -CUDART_INF_F, m3 = -CUDART_INF_F;
#pragma unroll
      for (int j = 0; j < KS_NK; j += 4) {
        if (j + 0 < kvalid) m0 = fmaxf(m0, __uint_as_float(sv[j + 0]));
        if (j + 1 < kvalid) m1 = fmaxf(m1, __uint_as_float(sv[j + 1]));
        if (j + 2 < kvalid) m2 = fmaxf(m2, __uint_as_float(sv[j + 2]));
        if (j + 3 < kvalid) m3 = fmaxf(m3, __uint_as_float(sv[j + 3]));
      }
      const float mc = fmaxf(fmaxf(m0, m1), fmaxf(m2, m3)) * c;
      // P buffer (i & 1) was last read by P V(i-2), whose completion the drain of O(i-2) has already waited for
      float l0 = 0.f, l1 = 0.f, l2 = 0.f, l3 = 0.f;
      const uint32_t pdst = tm_P + lane_off + (i & 1) * 40;
#pragma unroll
      for (int cc = 0; cc < KS_NK; cc += 16) {
        uint32_t pk[8];
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float e0 = ex2_approx(fmaf(__uint_as_float(sv[cc + j + 0]), c, -mc));
          float e1 = ex2_approx(fmaf(__uint_as_float(sv[cc + j + 1]), c, -mc));
          float e2 = ex2_approx(fmaf(__uint_as_float(sv[cc + j + 2]), c, -mc));
          float e3 = ex2_approx(fmaf(__uint_as_float(sv[cc + j + 3]), c, -mc));
          if (cc + j + 0 >= kvalid) e0 = 0.f;
          if (cc + j + 1 >= kvalid) e1 = 0.f;
          if (cc + j + 2 >= kvalid) e2 = 0.f;
          if (cc + j + 3 >= kvalid) e3 = 0.f;
          l0 += e0; l1 += e1; l2 += e2; l3 += e3;
          pk[(j >> 1) + 0] = pack_bf16(e0, e1);
          pk[(j >> 1) + 1] = pack_bf16(e2, e3);
        }
        asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};\n" ::"r"(pdst + (cc >> 1)),
                     "r"(pk[0]), "r"(pk[1]), "r"(pk[2]), "r"(pk[3]), "r"(pk[4]), "r"(pk[5]), "r"(pk[6]), "r"(pk[7])
                     : "memory");
      }
      inv_l[i & 1] = 1.f / ((l0 + l1) + (l2 + l3));
      tmem_st_wait();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(p_full(i & 1));
      if (i > 0) drain(i - 1);
    }
    drain(nt - 1);
    if (elect_one()) tma_store_wait_all<0>();   // bulk stores must have completed before the CTA releases its smem
    __syncwarp();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem, AT_TMEM_COLS);
  }
}


// 4-D view (column, token, inner sequence index, outer sequence index) of a row-major [rows, ld] bf16 matrix
int make_map(CUtensorMap* m, const void* base, int64_t ld, int cols, int L, int64_t n_inner, int64_t n_outer,
             int64_t tok_stride, int64_t inner_stride, int64_t outer_stride, uint32_t box_rows = 128) {
  const uint64_t dims[4] = {(uint64_t)cols, (uint64_t)L, (uint64_t)n_inner, (uint64_t)n_outer};
  // a size-1 dimension never advances; give it any legal (multiple of 16 B, non-zero) stride
  const uint64_t row = (uint64_t)ld * 2;
  const uint64_t s1 = (uint64_t)tok_stride * row;
  const uint64_t s2 = n_inner > 1 ? (uint64_t)inner_stride * row : s1 * (uint64_t)L;
  const uint64_t s3 = n_outer > 1 ? (uint64_t)outer_stride * row : (s2 > s1 ? s2 : s1) * 2;
  const uint64_t str[3] = {s1, s2, s3};
  const uint32_t box[4] = {AT_D, box_rows, 1, 1};
  return make_tmap_bf16(m, base, 4, dims, str, box, nullptr, 3);
}

int launch_attention(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* o,
                     int64_t ldo, int H, int Lq, int Lk, int64_t n_outer, int n_inner, int64_t q_outer, int64_t q_inner,
                     int64_t q_tok, int64_t kv_outer, int64_t kv_inner, int64_t kv_tok, float scale, int causal,
                     void* stream, int group = 0, const float* bias = nullptr, int bias_nw = 1) {
  AttnParams p{};
  p.causal = causal;
  p.group = group;
  p.bias = bias; p.bias_nw = bias_nw;
  p.tile_rows = group > 0 ? (AT_BM / group) * group : AT_BM;
  p.H = H; p.Lq = Lq; p.Lk = Lk; p.n_inner = n_inner;
  p.q_tiles = (Lq + p.tile_rows - 1) / p.tile_rows;
  p.nblk = group > 0 ? 1 : (Lk + AT_BN - 1) / AT_BN;
  p.scale_log2 = scale * 1.4426950408889634f;
  p.o = reinterpret_cast<__nv_bfloat16*>(o);
  p.ldo = ldo;
  p.o_outer = q_outer; p.o_inner = q_inner; p.o_tok = q_tok;
  CUtensorMap tmQ, tmK, tmV;
  int rc;
  if ((rc = make_map(&tmQ, q, ldq, H * 64, Lq, n_inner, n_outer, q_tok, q_inner, q_outer))) return rc;
  if ((rc = make_map(&tmK, k, ldk, H * 64, Lk, n_inner, n_outer, kv_tok, kv_inner, kv_outer))) return rc;
  if ((rc = make_map(&tmV, v, ldv, H * 64, Lk, n_inner, n_outer, kv_tok, kv_inner, kv_outer))) return rc;
  static int kvs_mode = -1;   // TAIR_ATTN_KVS=0 disables the key/value-stationary kernel (A/B probe)
  if (kvs_mode < 0) { const char* e = getenv("TAIR_ATTN_KVS"); kvs_mode = (e && atoi(e) == 0) ? 0 : 1; }
  if (kvs_mode && group == 0 && !causal && n_inner == 1 && Lk <= KS_NK && bias == nullptr) {
    // short context: stream the query tiles of a (batch, head) past resident K / V.  Chunks of query tiles per CTA so
    // that roughly two CTAs per SM are in flight.
    const long pairs = (long)n_outer * H;
    // query tiles per CTA: minimise waves x (tiles per CTA + ~1.5 tiles of set-up) with two CTAs resident per SM
    const long slots = 2L * num_sms();
    int chunks = 1;
    double best = 1e30;
    for (int c = 1; c <= p.q_tiles; ++c) {
      const int tpc = (p.q_tiles + c - 1) / c;
      const int cc = (p.q_tiles + tpc - 1) / tpc;
      const double cost = (double)((pairs * cc + slots - 1) / slots) * (tpc + 1.5);
      if (cost < best - 1e-9) { best = cost; chunks = cc; }
    }
    if (const char* e = getenv("TAIR_KVS_CHUNKS")) { const int v = atoi(e); if (v >= 1 && v <= p.q_tiles) chunks = v; }   // probe
    const int tiles_per_cta = (p.q_tiles + chunks - 1) / chunks;
    chunks = (p.q_tiles + tiles_per_cta - 1) / tiles_per_cta;
    CUtensorMap tmO;
    if ((rc = make_map(&tmO, o, ldo, H * 64, Lq, n_inner, n_outer, q_tok, q_inner, q_outer, 32))) return rc;
    TAIR_SMEM_OPTIN(attn_kvs_kernel, KS_SMEM);
    const long gridk = pairs * chunks;
    TAIR_REQUIRE(gridk < (1l << 31), "attention: grid too large");
    TAIR_LAUNCH((attn_kvs_kernel), (unsigned)gridk, AT_THREADS, KS_SMEM, static_cast<cudaStream_t>(stream), tmQ, tmK, tmV, tmO,
                p, tiles_per_cta, chunks);
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return check_launch("attn_kvs_kernel");
  }
  TAIR_SMEM_OPTIN(attn_tc_kernel, AT_SMEM);
  const long grid = (long)n_outer * n_inner * H * p.q_tiles;
  TAIR_REQUIRE(grid < (1l << 31), "attention: grid too large");
  TAIR_LAUNCH((attn_tc_kernel), (unsigned)grid, AT_THREADS, AT_SMEM, static_cast<cudaStream_t>(stream), tmQ, tmK, tmV, p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("attn_tc_kernel");
}

}  // namespace
}  // namespace tair

using namespace tair;

extern "C" int tair_attention_bf16(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v,
                                   int64_t ldv, void* o, int64_t ldo, int32_t B, int32_t H, int32_t Lq,
                                   int32_t Lk, int32_t head_dim, float scale, int32_t causal, void* stream) {
  TAIR_REQUIRE(q && k && v && o, "attention: NULL pointer");
  TAIR_REQUIRE(head_dim == 64, "attention: tensor-core path is built for head_dim 64 (got %d)", head_dim);
  TAIR_REQUIRE(B > 0 && H > 0 && Lq > 0 && Lk > 0, "attention: bad shape");
  TAIR_REQUIRE(ldq % 8 == 0 && ldk % 8 == 0 && ldv % 8 == 0 && ldo % 8 == 0,
               "attention: row strides must be multiples of 8 elements");
  TAIR_REQUIRE(ldq >= H * 64 && ldk >= H * 64 && ldv >= H * 64 && ldo >= H * 64,
               "attention: row stride smaller than H*head_dim");
  for (const void* ptr : {q, k, v, (const void*)o})
    TAIR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) % 16) == 0, "attention: pointers must be 16-byte aligned");
  TAIR_REQUIRE(!causal || Lq == Lk, "attention: causal masking needs Lq == Lk");
  return launch_attention(q, ldq, k, ldk, v, ldv, o, ldo, H, Lq, Lk, B, 1, Lq, 0, 1, Lk, 0, 1, scale, causal ? 1 : 0, stream);
}

extern "C" int tair_attention_seq_bf16(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo,
                                       int32_t H, int32_t L, int64_t n_outer, int32_t n_inner, int64_t outer_stride,
                                       int64_t inner_stride, int64_t tok_stride, float scale, void* stream) {
  TAIR_REQUIRE(q && k && v && o, "attention_seq: NULL pointer");
  TAIR_REQUIRE(H > 0 && L > 0 && n_outer > 0 && n_inner > 0 && tok_stride > 0, "attention_seq: bad shape");
  TAIR_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && ld >= H * 64 && ldo >= H * 64, "attention_seq: bad row strides");
  for (const void* ptr : {q, k, v, (const void*)o})
    TAIR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) % 16) == 0, "attention_seq: pointers must be 16-byte aligned");
  // Short sequences stored back to back (the TESTR decoder's intra-group attention: 100*B sequences of 16 / 25 points,
  // deformable_transformer.py:454-466) would each occupy a 128-row tensor-core tile for 16-25 useful rows; they are
  // packed 128/L to a tile instead and attended block-diagonally.
  if (n_inner == 1 && tok_stride == 1 && outer_stride == L && L <= 64 && n_outer * L < (1ll << 31))
    return launch_attention(q, ld, k, ld, v, ld, o, ldo, H, (int)(n_outer * L), (int)(n_outer * L), 1, 1, 0, 0, 1, 0, 0, 1,
                            scale, 0, stream, L);
  return launch_attention(q, ld, k, ld, v, ld, o, ldo, H, L, L, n_outer, n_inner, outer_stride, inner_stride,
                          tok_stride, outer_stride, inner_stride, tok_stride, scale, 0, stream);
}

extern "C" int tair_attention_windows_bf16(const void* q, const void* k, const void* v, int64_t ld, void* o, int64_t ldo,
                                           int32_t H, int32_t L, int64_t n_windows, const float* bias, int32_t bias_nw,
                                           float scale, void* stream) {
  TAIR_REQUIRE(q && k && v && o, "attention_windows: NULL pointer");
  TAIR_REQUIRE(H > 0 && L > 0 && L <= 64 && n_windows > 0 && n_windows * L < (1ll << 31), "attention_windows: bad shape");
  TAIR_REQUIRE(bias == nullptr || bias_nw > 0, "attention_windows: bias_nw must be positive");
  TAIR_REQUIRE(ld % 8 == 0 && ldo % 8 == 0 && ld >= H * 64 && ldo >= H * 64, "attention_windows: bad row strides");
  for (const void* ptr : {q, k, v, (const void*)o})
    TAIR_REQUIRE((reinterpret_cast<uintptr_t>(ptr) % 16) == 0, "attention_windows: pointers must be 16-byte aligned");
  return launch_attention(q, ld, k, ld, v, ld, o, ldo, H, (int)(n_windows * L), (int)(n_windows * L), 1, 1, 0, 0, 1, 0, 0, 1,
                          scale, 0, stream, L, bias, bias_nw);
}

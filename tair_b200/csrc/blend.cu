// Overlapped-tile blend + stitch as ONE kernel.
//
// Replaces merge_patches_with_overlap (val_patches.py:114-206): a Python loop of 2 slice-adds per tile
// plus a divide and a crop.  Here each output pixel gathers the (at most 2 x 2) tiles that cover it.
//   window(l) = min(1, (l+1)/fade, (T-l)/fade)  per axis, 2-D window = product      (:155-167)
//   out = sum_i w_i * tile_i / max(sum_i w_i, 1e-8), accumulated in the reference's tile order
//   (row-major i, j) with unfused fp32 multiply / add, so the result is bit-identical.
// Canvas = (n_h-1)*stride + T square-ish grid; only the top-left out_h x out_w crop is produced (:200-204).
#include <atomic>

#include "../../include/tair_b200.h"
#include "common.cuh"

namespace tair {
extern std::atomic<int64_t> g_launch_count;
namespace {

struct BlendParams {
  const float* tiles;  // [n_tiles, C, T, T]
  float* out;          // [C, out_h, out_w]
  int n_tiles, n_h, n_w, C, T, stride, fade, out_h, out_w;
};

__device__ __forceinline__ float ramp(int l, int T, int fade) {
  // the reference forms (i+1)/fade as a Python double and rounds it once to fp32
  if (l < fade) return (float)((double)(l + 1) / (double)fade);
  if (l >= T - fade) return (float)((double)(T - l) / (double)fade);
  return 1.f;
}

__global__ void blend_tiles_kernel(const BlendParams p) {
  pdl_grid_sync();
  const int x = blockIdx.x * blockDim.x + threadIdx.x;
  const int y = blockIdx.y;
  if (x >= p.out_w) return;
  int i_lo;  // smallest i with i*stride + T > y
  if (y - p.T + 1 <= 0) i_lo = 0; else i_lo = (y - p.T + p.stride) / p.stride;
  int i_hi = y / p.stride;
  if (i_hi > p.n_h - 1) i_hi = p.n_h - 1;
  int j_lo;
  if (x - p.T + 1 <= 0) j_lo = 0; else j_lo = (x - p.T + p.stride) / p.stride;
  int j_hi = x / p.stride;
  if (j_hi > p.n_w - 1) j_hi = p.n_w - 1;
  for (int c = 0; c < p.C; ++c) {
    float acc = 0.f, wsum = 0.f;
    for (int i = i_lo; i <= i_hi; ++i) {
      const int ly = y - i * p.stride;
      const float wy = ramp(ly, p.T, p.fade);
      for (int j = j_lo; j <= j_hi; ++j) {
        const int t = i * p.n_w + j;
        if (t >= p.n_tiles) continue;  // reference stops placing tiles when the list runs out (:171-173)
        const int lx = x - j * p.stride;
        const float w = __fmul_rn(wy, ramp(lx, p.T, p.fade));
        const float v = __ldg(p.tiles + (((long)t * p.C + c) * p.T + ly) * p.T + lx);
        acc = __fadd_rn(acc, __fmul_rn(v, w));
        wsum = __fadd_rn(wsum, w);
      }
    }
    wsum = fmaxf(wsum, 1e-8f);
    p.out[((long)c * p.out_h + y) * p.out_w + x] = __fdiv_rn(acc, wsum);
  }
}

}  // namespace
}  // namespace tair

using namespace tair;

extern "C" int tair_blend_tiles(const float* tiles, float* out, int32_t n_tiles, int32_t n_h, int32_t n_w,
                                int32_t C, int32_t tile, int32_t overlap, int32_t out_h, int32_t out_w,
                                void* stream) {
  TAIR_REQUIRE(tiles && out, "blend_tiles: NULL pointer");
  TAIR_REQUIRE(n_tiles > 0 && n_h > 0 && n_w > 0 && C > 0 && tile > 0, "blend_tiles: bad shape");
  TAIR_REQUIRE(overlap >= 0 && 2 * overlap <= tile, "blend_tiles: overlap must satisfy 0 <= 2*overlap <= tile");
  const int stride = tile - overlap;
  TAIR_REQUIRE(out_h > 0 && out_w > 0 && out_h <= (n_h - 1) * stride + tile && out_w <= (n_w - 1) * stride + tile,
               "blend_tiles: output crop (%d x %d) exceeds the tile canvas", out_h, out_w);
  BlendParams p{tiles, out, n_tiles, n_h, n_w, C, tile, stride, overlap > 0 ? overlap : 1, out_h, out_w};
  if (overlap == 0) p.fade = 0;
  dim3 block(256), grid((out_w + 255) / 256, out_h);
  TAIR_LAUNCH((blend_tiles_kernel), grid, block, 0, static_cast<cudaStream_t>(stream), p);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("blend_tiles_kernel");
}

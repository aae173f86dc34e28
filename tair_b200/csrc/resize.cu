// Tile front-end on the GPU: crop 128x128 LQ tiles out of the zero-padded image and resize them to 512x512 with the
// SAME arithmetic as PIL's Image.resize(..., BICUBIC) on 8-bit images, then scale to [0,1] fp32 planar — the
// per-tile `preprocess_lq` of the reference (val_patches.py:291-294,318: T.Resize(BICUBIC) + T.ToTensor()).
//
// PIL resamples in two passes, horizontal then vertical, each in 22-bit fixed point with an 8-bit intermediate:
//     out = clip8((2^21 + sum_k pixel[xmin + k] * coeff[k]) >> 22)
// with per-output-index windows [xmin, xmin + n) and integer coefficients that the host computes exactly as
// Resample.c does (tair_b200/tiles.py: pil_bicubic_coeffs).  Both passes are reproduced bit for bit, so the GPU
// front-end is a drop-in for the host one (tests compare against PIL itself).
#include <atomic>

#include "../../include/tair_b200.h"
#include "common.cuh"

namespace tair {
extern std::atomic<int64_t> g_launch_count;
namespace {

__device__ __forceinline__ int clip8(int v) {
  v >>= 22;
  return v < 0 ? 0 : (v > 255 ? 255 : v);
}

// pass 1: [P tiles][tile rows][out cols][3] u8  <-  image [Hp][Wp][3] u8
__global__ void resize_h_kernel(const uint8_t* __restrict__ img, int Wp, const int32_t* __restrict__ origins,
                                const int32_t* __restrict__ bounds, const int32_t* __restrict__ coeffs, int ksize,
                                int tile, int out, uint8_t* __restrict__ tmp, long total) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % out);
  const long t = idx / out;
  const int row = (int)(t % tile);
  const int p = (int)(t / tile);
  const int y0 = origins[2 * p], x0 = origins[2 * p + 1];
  const int xmin = bounds[2 * ox], n = bounds[2 * ox + 1];
  const uint8_t* src = img + ((long)(y0 + row) * Wp + x0 + xmin) * 3;
  int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
  for (int k = 0; k < n; ++k) {
    const int c = coeffs[ox * ksize + k];
    s0 += src[3 * k] * c; s1 += src[3 * k + 1] * c; s2 += src[3 * k + 2] * c;
  }
  uint8_t* dst = tmp + idx * 3;
  dst[0] = (uint8_t)clip8(s0); dst[1] = (uint8_t)clip8(s1); dst[2] = (uint8_t)clip8(s2);
}

// pass 2: [P][3][out][out] fp32 in [0,1]  <-  tmp [P][tile][out][3] u8
__global__ void resize_v_kernel(const uint8_t* __restrict__ tmp, const int32_t* __restrict__ bounds,
                                const int32_t* __restrict__ coeffs, int ksize, int tile, int out,
                                float* __restrict__ dst, long total) {
  pdl_grid_sync();
  const long idx = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int ox = (int)(idx % out);
  const long t = idx / out;
  const int oy = (int)(t % out);
  const int p = (int)(t / out);
  const int ymin = bounds[2 * oy], n = bounds[2 * oy + 1];
  const uint8_t* src = tmp + (((long)p * tile + ymin) * out + ox) * 3;
  int s0 = 1 << 21, s1 = 1 << 21, s2 = 1 << 21;
  for (int k = 0; k < n; ++k) {
    const int c = coeffs[oy * ksize + k];
    const uint8_t* q = src + (long)k * out * 3;
    s0 += q[0] * c; s1 += q[1] * c; s2 += q[2] * c;
  }
  const long plane = (long)out * out;
  float* o = dst + (long)p * 3 * plane + (long)oy * out + ox;
  o[0] = __fdiv_rn((float)clip8(s0), 255.f);           // ToTensor(): uint8 / 255 in fp32
  o[plane] = __fdiv_rn((float)clip8(s1), 255.f);
  o[2 * plane] = __fdiv_rn((float)clip8(s2), 255.f);
}

}  // namespace
}  // namespace tair

using namespace tair;

extern "C" int tair_tiles_bicubic_u8(const void* image, int32_t Hp, int32_t Wp, const int32_t* origins, int32_t P,
                                     int32_t tile, int32_t out, const int32_t* bounds, const int32_t* coeffs,
                                     int32_t ksize, void* tmp, float* dst, void* stream) {
  TAIR_REQUIRE(image && origins && bounds && coeffs && tmp && dst, "tiles_bicubic: NULL pointer");
  TAIR_REQUIRE(Hp > 0 && Wp > 0 && P > 0 && tile > 0 && out > 0 && ksize > 0 && tile <= Hp && tile <= Wp,
               "tiles_bicubic: bad shape");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long t1 = (long)P * tile * out, t2 = (long)P * out * out;
  TAIR_LAUNCH((resize_h_kernel), (unsigned)((t1 + 255) / 256), 256, 0, st, reinterpret_cast<const uint8_t*>(image), Wp, origins,
                                                                 bounds, coeffs, ksize, tile, out,
                                                                 reinterpret_cast<uint8_t*>(tmp), t1);
  int rc = check_launch("resize_h_kernel");
  if (rc) return rc;
  TAIR_LAUNCH((resize_v_kernel), (unsigned)((t2 + 255) / 256), 256, 0, st, reinterpret_cast<const uint8_t*>(tmp), bounds, coeffs,
                                                                 ksize, tile, out, dst, t2);
  g_launch_count.fetch_add(2, std::memory_order_relaxed);
  return check_launch("resize_v_kernel");
}

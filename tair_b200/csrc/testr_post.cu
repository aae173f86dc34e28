// Detection post-processing of the TESTR head in ONE kernel (testr/adet/modeling/transformer_detector.py:123-152 and the
// per-instance host loop of terediff/sampler/spaced_sampler.py:298-306):
//   score   = sigmoid(mean over the 16 control points of the point-class logit)         (num_classes == 1)
//   polygon = control points scaled to pixels: x * image_w, y * image_h
//   recs    = arg-max character per position (arg-max of the logits == arg-max of softmax(logits))
// The reference runs softmax / mean / sigmoid / max / boolean-mask indexing / topk as ~15 eager launches plus one
// device->host copy per instance; with B tiles x 100 queries in flight that was 15 ms of a 48 ms step.  Here every (tile,
// query) is handled by one warp, results land in compact arrays that go to the host in a single copy, and the score
// threshold is applied there.
#include <atomic>

#include "../../include/tair_b200.h"
#include "common.cuh"

namespace tair {
extern std::atomic<int64_t> g_launch_count;
namespace {

__global__ void __launch_bounds__(256)
testr_post_kernel(const float* __restrict__ logits, const float* __restrict__ coords, const float* __restrict__ texts,
                  float* __restrict__ scores, float* __restrict__ polys, uint8_t* __restrict__ recs, int n_items,
                  int n_pts, int n_chars, int voc, float img_w, float img_h) {
  pdl_grid_sync();
  const int item = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (item >= n_items) return;
  // score: fixed-order sum of the point logits (n_pts <= 32), as torch.mean over a contiguous axis
  float v = lane < n_pts ? logits[(int64_t)item * n_pts + lane] : 0.f;
  float s = 0.f;
  for (int i = 0; i < n_pts; ++i) s += __shfl_sync(0xffffffffu, v, i);
  if (lane == 0) scores[item] = 1.f / (1.f + expf(-(s / (float)n_pts)));
  // polygon in pixels: coordinate pairs (x, y)
  for (int i = lane; i < 2 * n_pts; i += 32)
    polys[(int64_t)item * 2 * n_pts + i] = coords[(int64_t)item * 2 * n_pts + i] * ((i & 1) ? img_h : img_w);
  // characters: first arg-max over the vocabulary per position
  const float* tp = texts + (int64_t)item * n_chars * voc;
  for (int c = 0; c < n_chars; ++c) {
    float best = -INFINITY;
    int bi = 0x7fffffff;
    for (int k = lane; k < voc; k += 32) {
      const float x = tp[c * voc + k];
      if (x > best) { best = x; bi = k; }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob > best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) recs[(int64_t)item * n_chars + c] = (uint8_t)bi;
  }
}

}  // namespace
}  // namespace tair

using namespace tair;

extern "C" int tair_testr_postprocess(const float* pred_logits, const float* pred_ctrl_points, const float* pred_texts,
                                      float* scores, float* polygons, uint8_t* recs, int32_t n_items, int32_t n_pts,
                                      int32_t n_chars, int32_t voc, float image_w, float image_h, void* stream) {
  TAIR_REQUIRE(pred_logits && pred_ctrl_points && pred_texts && scores && polygons && recs, "testr_postprocess: NULL pointer");
  TAIR_REQUIRE(n_items > 0 && n_pts > 0 && n_pts <= 32 && n_chars > 0 && voc > 0 && voc <= 256,
               "testr_postprocess: needs 0 < n_pts <= 32 and 0 < voc <= 256 (n_pts=%d voc=%d)", n_pts, voc);
  const int warps = 8;
  TAIR_LAUNCH((testr_post_kernel), (n_items + warps - 1) / warps, warps * 32, 0, static_cast<cudaStream_t>(stream),
              pred_logits, pred_ctrl_points, pred_texts, scores, polygons, recs, n_items, n_pts, n_chars, voc, image_w,
              image_h);
  g_launch_count.fetch_add(1, std::memory_order_relaxed);
  return check_launch("testr_post_kernel");
}

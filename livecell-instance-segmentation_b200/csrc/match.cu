// match.cu — training-side siblings of the region path.
//
//   * box IoU matrix and the fused row max / argmax.  Replaces torchvision.ops.box_iou
//     (TV:ops/boxes.py:308-370: lt/rb broadcast -> [N,G,2] temporaries -> clamp -> inter -> union ->
//     divide, ~8 ATen launches) and the `ious.max(dim=1)` that follows it at
//     src/components/rpn.py:72-73, src/custom_maskrcnn.py:221-222,249-250, src/utils/mask_utils.py:93-94.
//     The fused kernel never materialises the [N,G] matrix: 16 B in + 12 B out per anchor.
//   * mask-target extraction, batched.  Replaces the per-positive Python loop over
//     extract_mask_target (src/utils/mask_utils.py:6-46, called at :110-113): int-box crop of the
//     matched uint8 ground-truth mask + bilinear resize (align_corners=False) to M x M, with two
//     `.item()` host syncs per proposal in the reference.
//
// IoU arithmetic is torchvision's, operation for operation (fp32, no contraction):
//   area = (x2-x1)*(y2-y1); wh = max(min(rb) - max(lt), 0); inter = w*h; iou = inter/((a1+a2)-inter).
// Row max follows torch.max(dim=1): first index of the maximum, NaN (0/0: two zero-area boxes)
// propagates and wins.
#include "common.cuh"

namespace lcr {

constexpr int kIouThreads = 256;
constexpr int kGtChunk = 1024;

__device__ __forceinline__ float iou_tv(const float4 a, float area_a, const float4 b, float area_b) {
  const float left = fmaxf(a.x, b.x), right = fminf(a.z, b.z);
  const float top = fmaxf(a.y, b.y), bottom = fminf(a.w, b.w);
  const float w = fmaxf(__fsub_rn(right, left), 0.f), h = fmaxf(__fsub_rn(bottom, top), 0.f);
  const float inter = __fmul_rn(w, h);
  const float uni = __fsub_rn(__fadd_rn(area_a, area_b), inter);
  if (inter == 0.f && uni > 0.f) return 0.f;  // the common case: no IEEE division
  return __fdiv_rn(inter, uni);
}

__device__ __forceinline__ float box_area_tv(const float4 b) { return __fmul_rn(__fsub_rn(b.z, b.x), __fsub_rn(b.w, b.y)); }

// One thread per row box; ground-truth boxes staged through shared memory in chunks.
template <bool FULL>
__global__ void __launch_bounds__(kIouThreads) iou_kernel(const float4* __restrict__ boxes, int N, const float4* __restrict__ gt, int G,
                                                          float* __restrict__ iou_out, float* __restrict__ max_out,
                                                          long long* __restrict__ arg_out) {
  __shared__ float4 s_gt[kGtChunk];
  __shared__ float s_area[kGtChunk];
  const int i = blockIdx.x * kIouThreads + threadIdx.x;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < N) a = __ldg(boxes + i);
  const float area_a = box_area_tv(a);
  float best = 0.f;
  int best_j = 0;
  bool best_nan = false, have = false;
  for (int g0 = 0; g0 < G; g0 += kGtChunk) {
    const int ng = min(kGtChunk, G - g0);
    __syncthreads();
    for (int j = threadIdx.x; j < ng; j += kIouThreads) {
      const float4 b = __ldg(gt + g0 + j);
      s_gt[j] = b;
      s_area[j] = box_area_tv(b);
    }
    __syncthreads();
    if (i < N) {
      for (int j = 0; j < ng; ++j) {
        const float v = iou_tv(a, area_a, s_gt[j], s_area[j]);
        if (FULL) {
          iou_out[(size_t)i * G + g0 + j] = v;
        } else if (!best_nan && (!have || v > best || v != v)) {
          best = v;
          best_j = g0 + j;
          best_nan = (v != v);
          have = true;
        }
      }
    }
  }
  if (!FULL && i < N) {
    max_out[i] = best;
    arg_out[i] = (long long)best_j;
  }
}

// Fused anchor / proposal matcher: IoU row max + the two threshold masks + their population counts, one pass.
// Follows src/components/rpn.py:72-81 (`pos_mask = max_ious >= 0.5; neg_mask = max_ious < 0.3; pos_mask.sum(); neg_mask.sum()`)
// and src/custom_maskrcnn.py:221-225, 249-251 (`max_iou >= 0.4`): five ATen launches over an [N,G] matrix become one kernel
// reading 16 B and writing 14 B per row box.  The row loop is iou_kernel<false>'s, statement for statement, so max / argmax
// are the same bits; the comparisons are fp32 against fp32 thresholds (what ATen's compare-with-scalar does for a float
// tensor); a NaN row (two zero-area boxes) is neither positive nor negative, as in the reference.
// Measured and dropped (profiles/r02d_match_ncu.md): a short path that skips iou_tv for pairs with an empty intersection and
// a finite positive area sum (bit-identical on the hostile-input test) is SLOWER — 62 us against 52 us at C1 size: lanes of
// a warp split between the two paths, and iou_tv's own `inter == 0` shortcut already keeps the division off the common path.
__global__ void __launch_bounds__(kIouThreads) match_kernel(const float4* __restrict__ boxes, int N, const float4* __restrict__ gt, int G,
                                                            float pos_thr, float neg_thr, float* __restrict__ max_out,
                                                            long long* __restrict__ arg_out, uint8_t* __restrict__ pos_out,
                                                            uint8_t* __restrict__ neg_out, int* __restrict__ counts) {
  __shared__ float4 s_gt[kGtChunk];
  __shared__ float s_area[kGtChunk];
  __shared__ int s_cnt[2];
  if (threadIdx.x < 2) s_cnt[threadIdx.x] = 0;
  const int i = blockIdx.x * kIouThreads + threadIdx.x;
  float4 a = make_float4(0.f, 0.f, 0.f, 0.f);
  if (i < N) a = __ldg(boxes + i);
  const float area_a = box_area_tv(a);
  float best = 0.f;
  int best_j = 0;
  bool best_nan = false, have = false;
  for (int g0 = 0; g0 < G; g0 += kGtChunk) {
    const int ng = min(kGtChunk, G - g0);
    __syncthreads();
    for (int j = threadIdx.x; j < ng; j += kIouThreads) {
      const float4 b = __ldg(gt + g0 + j);
      s_gt[j] = b;
      s_area[j] = box_area_tv(b);
    }
    __syncthreads();
    if (i < N) {
      for (int j = 0; j < ng; ++j) {
        const float v = iou_tv(a, area_a, s_gt[j], s_area[j]);
        if (!best_nan && (!have || v > best || v != v)) {
          best = v;
          best_j = g0 + j;
          best_nan = (v != v);
          have = true;
        }
      }
    }
  }
  const bool pos = (i < N) && (best >= pos_thr);  // false for NaN
  const bool neg = (i < N) && (best < neg_thr);
  if (i < N) {
    if (max_out) max_out[i] = best;
    if (arg_out) arg_out[i] = (long long)best_j;
    pos_out[i] = pos ? 1 : 0;
    neg_out[i] = neg ? 1 : 0;
  }
  // counts: one ballot per warp, one shared add per warp, one global add per CTA and mask
  const unsigned pb = __ballot_sync(0xffffffffu, pos), nb = __ballot_sync(0xffffffffu, neg);
  if (lane_id() == 0) {
    if (pb) atomicAdd(&s_cnt[0], __popc(pb));
    if (nb) atomicAdd(&s_cnt[1], __popc(nb));
  }
  __syncthreads();
  if (threadIdx.x < 2 && s_cnt[threadIdx.x]) atomicAdd(counts + threadIdx.x, s_cnt[threadIdx.x]);
}

// ATen upsample_bilinear2d source index (align_corners=False), see paste.cu / SURVEY App. B.4.
__device__ __forceinline__ void src_index_mt(float scale, int dst, int in_size, int& i0, int& i1, float& l0, float& l1) {
  float src = __fmaf_rn(scale, (float)dst + 0.5f, -0.5f);
  src = src < 0.f ? 0.f : src;
  i0 = min(__float2int_rz(src), in_size - 1);
  i1 = i0 + ((i0 < in_size - 1) ? 1 : 0);
  l1 = __fsub_rn(src, (float)i0);
  l0 = __fsub_rn(1.0f, l1);
}

// One CTA per target: M*M outputs, each a bilinear tap of the cropped uint8 mask.
__global__ void __launch_bounds__(256) mask_targets_kernel(const uint8_t* __restrict__ gt_masks, int G, int H, int W,
                                                           const float4* __restrict__ boxes, const long long* __restrict__ gt_index,
                                                           int K, int M, float* __restrict__ out) {
  const int k = blockIdx.x;
  const long long gi = gt_index ? gt_index[k] : (long long)k;
  float* o = out + (size_t)k * M * M;
  if (gi < 0 || gi >= G) {  // unmatched slot: zeros
    for (int i = threadIdx.x; i < M * M; i += blockDim.x) o[i] = 0.f;
    return;
  }
  const float4 b = __ldg(boxes + k);
  // x1, y1, x2, y2 = box.int(); clip exactly as src/utils/mask_utils.py:23-31
  int x1 = __float2int_rz(b.x), y1 = __float2int_rz(b.y), x2 = __float2int_rz(b.z), y2 = __float2int_rz(b.w);
  x1 = max(0, min(x1, W - 1));
  y1 = max(0, min(y1, H - 1));
  x2 = max(x1 + 1, min(x2, W));
  y2 = max(y1 + 1, min(y2, H));
  const int ch = y2 - y1, cw = x2 - x1;
  const float sh = __fdiv_rn((float)ch, (float)M), sw = __fdiv_rn((float)cw, (float)M);
  const uint8_t* m = gt_masks + (size_t)gi * H * W + (size_t)y1 * W + x1;
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) {
    const int oy = i / M, ox = i - oy * M;
    int h0, h1, w0, w1;
    float wy0, wy1, wx0, wx1;
    src_index_mt(sh, oy, ch, h0, h1, wy0, wy1);
    src_index_mt(sw, ox, cw, w0, w1, wx0, wx1);
    const float a = (float)m[(size_t)h0 * W + w0], bq = (float)m[(size_t)h0 * W + w1];
    const float c = (float)m[(size_t)h1 * W + w0], d = (float)m[(size_t)h1 * W + w1];
    const float top = __fmaf_rn(a, wx0, __fmul_rn(bq, wx1));
    const float bot = __fmaf_rn(c, wx0, __fmul_rn(d, wx1));
    o[i] = __fmaf_rn(top, wy0, __fmul_rn(bot, wy1));
  }
}

// ---- mask-head tail -----------------------------------------------------------
// CustomMaskHead.forward ends with F.interpolate(mask_logits [K,2,m,m] -> [K,2,M,M], bilinear,
// align_corners=False) (src/components/mask_head.py:52-58) and _generate_masks then takes
// sigmoid(mask_logits[:, 1]) (src/custom_maskrcnn.py:273-274): three passes over [K,2,28,28] of which half
// is thrown away.  One kernel: class `cls` only, upsample + sigmoid fused, output = the paste kernel's input.
__global__ void __launch_bounds__(256) mask_tail_kernel(const float* __restrict__ logits, int K, int Cc, int cls, int m, int M,
                                                        float* __restrict__ probs) {
  const int k = blockIdx.x;
  const float* src = logits + ((size_t)k * Cc + cls) * m * m;
  float* o = probs + (size_t)k * M * M;
  const float scale = __fdiv_rn((float)m, (float)M);
  for (int i = threadIdx.x; i < M * M; i += blockDim.x) {
    const int oy = i / M, ox = i - oy * M;
    float v;
    if (m == M) {
      v = __ldg(src + i);
    } else {
      int h0, h1, w0, w1;
      float wy0, wy1, wx0, wx1;
      src_index_mt(scale, oy, m, h0, h1, wy0, wy1);
      src_index_mt(scale, ox, m, w0, w1, wx0, wx1);
      const float a = __ldg(src + h0 * m + w0), b = __ldg(src + h0 * m + w1);
      const float c = __ldg(src + h1 * m + w0), d = __ldg(src + h1 * m + w1);
      const float top = __fmaf_rn(a, wx0, __fmul_rn(b, wx1));
      const float bot = __fmaf_rn(c, wx0, __fmul_rn(d, wx1));
      v = __fmaf_rn(top, wy0, __fmul_rn(bot, wy1));
    }
    o[i] = sigmoid_f32(v);
  }
}

// ---- tile stitching -----------------------------------------------------------
// calculate_mask_area_in_region (src/visualize.py:106-130) as used by
// filter_detections_by_border_mini_tiles (:174-257): per detection, the number of mask pixels inside each
// of its valid mini-tile rectangles and in the whole mask.  Integer counts out (exact); the float64
// fractions and the threshold stay on the host exactly as the reference computes them.
// One CTA per detection; when boxes are given only the box rows/columns are scanned (a pasted mask is zero
// outside its box).
constexpr int kMaxRects = 16;

__global__ void __launch_bounds__(256) mask_region_counts_kernel(const uint8_t* __restrict__ masks, int N, int H, int W,
                                                                 const float4* __restrict__ boxes, const int4* __restrict__ rects,
                                                                 const int* __restrict__ rect_offsets, int thr,
                                                                 int* __restrict__ total, int* __restrict__ in_region) {
  __shared__ int s_cnt[kMaxRects + 1];
  __shared__ int4 s_rect[kMaxRects];
  const int i = blockIdx.x;
  const int r0 = rect_offsets[i], nr_all = max(rect_offsets[i + 1] - r0, 0);
  int x1 = 0, y1 = 0, x2 = W, y2 = H;
  if (boxes) {  // same integer box as the paste kernel wrote into (src/custom_maskrcnn.py:279-283)
    const float4 b = __ldg(boxes + i);
    x1 = max(0, __float2int_rz(b.x));
    y1 = max(0, __float2int_rz(b.y));
    x2 = min(W, __float2int_rz(b.z));
    y2 = min(H, __float2int_rz(b.w));
  }
  const uint8_t* mk = masks + (size_t)i * H * W;
  const int bw = max(x2 - x1, 0), bh = max(y2 - y1, 0);
  // kMaxRects rectangles per pass over the box; a detection with more (the 25-tile stitcher needs at most 9) takes further
  // passes instead of being truncated
  for (int base = 0; base == 0 || base < nr_all; base += kMaxRects) {
    const int nr = min(nr_all - base, kMaxRects);
    __syncthreads();
    if (threadIdx.x <= kMaxRects) s_cnt[threadIdx.x] = 0;
    if (threadIdx.x < nr) s_rect[threadIdx.x] = rects[r0 + base + threadIdx.x];
    __syncthreads();
    int cnt[kMaxRects + 1];
#pragma unroll
    for (int r = 0; r <= kMaxRects; ++r) cnt[r] = 0;
    for (int p = threadIdx.x; p < bw * bh; p += blockDim.x) {
      const int y = y1 + p / bw, x = x1 + p % bw;
      if ((int)mk[(size_t)y * W + x] > thr) {
        cnt[0]++;
#pragma unroll
        for (int r = 0; r < kMaxRects; ++r)
          if (r < nr && x >= s_rect[r].x && x < s_rect[r].z && y >= s_rect[r].y && y < s_rect[r].w) cnt[r + 1]++;
      }
    }
#pragma unroll
    for (int r = 0; r <= kMaxRects; ++r) {
      const int v = __reduce_add_sync(0xFFFFFFFFu, cnt[r]);
      if (lane_id() == 0 && v) atomicAdd(&s_cnt[r], v);
    }
    __syncthreads();
    if (threadIdx.x == 0 && base == 0) total[i] = s_cnt[0];
    if (threadIdx.x < nr) in_region[r0 + base + threadIdx.x] = s_cnt[threadIdx.x + 1];
  }
}

}  // namespace lcr

using namespace lcr;

extern "C" int lcr_box_iou_f32(const float* boxes, int N, const float* gt, int G, float* iou, void* stream) {
  LCR_REQUIRE(N >= 0 && G >= 0, LCR_ERR_INVALID_ARG);
  if (N == 0 || G == 0) return LCR_OK;
  LCR_REQUIRE(boxes && gt && iou, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(boxes, 16) && aligned_to(gt, 16), LCR_ERR_ALIGNMENT);
  iou_kernel<true><<<(N + kIouThreads - 1) / kIouThreads, kIouThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(boxes), N, reinterpret_cast<const float4*>(gt), G, iou, nullptr, nullptr);
  return after_launch();
}

extern "C" int lcr_box_iou_max_f32(const float* boxes, int N, const float* gt, int G, float* max_iou, int64_t* argmax,
                                   void* stream) {
  LCR_REQUIRE(N >= 0 && G > 0, LCR_ERR_INVALID_ARG);  // torch.max over an empty dimension is an error too
  if (N == 0) return LCR_OK;
  LCR_REQUIRE(boxes && gt && max_iou && argmax, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(boxes, 16) && aligned_to(gt, 16), LCR_ERR_ALIGNMENT);
  iou_kernel<false><<<(N + kIouThreads - 1) / kIouThreads, kIouThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(boxes), N, reinterpret_cast<const float4*>(gt), G, nullptr, max_iou,
      reinterpret_cast<long long*>(argmax));
  return after_launch();
}

extern "C" int lcr_match_boxes_f32(const float* boxes, int N, const float* gt, int G, float pos_thr, float neg_thr, float* max_iou,
                                  int64_t* argmax, uint8_t* pos_mask, uint8_t* neg_mask, int* counts, void* stream) {
  LCR_REQUIRE(N >= 0 && G > 0, LCR_ERR_INVALID_ARG);  // the reference returns its no-ground-truth loss before matching
  LCR_REQUIRE(counts != nullptr, LCR_ERR_INVALID_ARG);
  {
    const int rc = cuda_status(cudaMemsetAsync(counts, 0, 2 * sizeof(int), as_stream(stream)));
    if (rc != LCR_OK) return rc;
  }
  if (N == 0) return LCR_OK;
  LCR_REQUIRE(boxes && gt && pos_mask && neg_mask, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(boxes, 16) && aligned_to(gt, 16), LCR_ERR_ALIGNMENT);
  match_kernel<<<(N + kIouThreads - 1) / kIouThreads, kIouThreads, 0, as_stream(stream)>>>(
      reinterpret_cast<const float4*>(boxes), N, reinterpret_cast<const float4*>(gt), G, pos_thr, neg_thr, max_iou,
      reinterpret_cast<long long*>(argmax), pos_mask, neg_mask, counts);
  return after_launch();
}

extern "C" int lcr_mask_targets_f32(const uint8_t* gt_masks, int G, int H, int W, const float* boxes, const int64_t* gt_index,
                                    int K, int M, float* out, void* stream) {
  LCR_REQUIRE(K >= 0 && G >= 0 && H > 0 && W > 0 && M > 0, LCR_ERR_INVALID_ARG);
  if (K == 0) return LCR_OK;
  LCR_REQUIRE(gt_masks && boxes && out && G > 0, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(boxes, 16), LCR_ERR_ALIGNMENT);
  LCR_REQUIRE((int64_t)H * W < (1ll << 31), LCR_ERR_CAPACITY);
  mask_targets_kernel<<<K, 256, 0, as_stream(stream)>>>(gt_masks, G, H, W, reinterpret_cast<const float4*>(boxes),
                                                       reinterpret_cast<const long long*>(gt_index), K, M, out);
  return after_launch();
}

extern "C" int lcr_mask_tail_f32(const float* logits, int K, int num_classes, int cls, int m, int M, float* probs, void* stream) {
  LCR_REQUIRE(K >= 0 && num_classes > 0 && cls >= 0 && cls < num_classes && m > 0 && M > 0, LCR_ERR_INVALID_ARG);
  if (K == 0) return LCR_OK;
  LCR_REQUIRE(logits && probs, LCR_ERR_INVALID_ARG);
  mask_tail_kernel<<<K, 256, 0, as_stream(stream)>>>(logits, K, num_classes, cls, m, M, probs);
  return after_launch();
}

extern "C" int lcr_mask_region_counts_u8(const uint8_t* masks, int N, int H, int W, const float* boxes, const int* rects,
                                         const int* rect_offsets, int threshold, int* total, int* in_region, void* stream) {
  LCR_REQUIRE(N >= 0 && H > 0 && W > 0, LCR_ERR_INVALID_ARG);
  if (N == 0) return LCR_OK;
  LCR_REQUIRE(masks && rect_offsets && total, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(!boxes || aligned_to(boxes, 16), LCR_ERR_ALIGNMENT);
  LCR_REQUIRE(!rects || aligned_to(rects, 16), LCR_ERR_ALIGNMENT);
  LCR_REQUIRE((int64_t)H * W < (1ll << 31), LCR_ERR_CAPACITY);
  mask_region_counts_kernel<<<N, 256, 0, as_stream(stream)>>>(masks, N, H, W, reinterpret_cast<const float4*>(boxes),
                                                             reinterpret_cast<const int4*>(rects), rect_offsets, threshold, total,
                                                             in_region);
  return after_launch();
}

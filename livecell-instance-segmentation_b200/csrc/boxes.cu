// boxes.cu — small box-side kernels: anchors, clip, min-size mask, delta decode, FPN level map,
// keep-list gather, detection records.  All are one-thread-per-box, float4-vectorised, coalesced.
#include "common.cuh"

namespace lcr {

struct BaseAnchors {
  float v[LCR_MAX_ANCHORS * 4];
};

// a1 — AnchorGenerator.generate_anchors (src/components/anchor_generator.py:29-35):
// anchors[(y*w+x)*A + a] = fp32(x*stride, y*stride, x*stride, y*stride) + base[a].
__global__ void __launch_bounds__(256) anchors_kernel(float4* __restrict__ out, int h, int w, int A, float stride,
                                                       const __grid_constant__ BaseAnchors base) {
  const int n = h * w * A;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    const int a = i % A, p = i / A;
    const int y = p / w, x = p - y * w;
    const float sx = __fmul_rn((float)x, stride), sy = __fmul_rn((float)y, stride);
    float4 o;
    o.x = __fadd_rn(sx, base.v[a * 4 + 0]);
    o.y = __fadd_rn(sy, base.v[a * 4 + 1]);
    o.z = __fadd_rn(sx, base.v[a * 4 + 2]);
    o.w = __fadd_rn(sy, base.v[a * 4 + 3]);
    out[i] = o;
  }
}

// a5 — clip_boxes_to_image (src/utils/box_utils.py:32-37), in place.
__global__ void __launch_bounds__(256) clip_kernel(float4* __restrict__ boxes, int K, float img_h, float img_w) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  float4 b = boxes[i];
  b.x = clampf(b.x, 0.f, img_w);
  b.z = clampf(b.z, 0.f, img_w);
  b.y = clampf(b.y, 0.f, img_h);
  b.w = clampf(b.w, 0.f, img_h);
  boxes[i] = b;
}

// a6 — filter_small_boxes (src/utils/box_utils.py:39-44).
__global__ void __launch_bounds__(256) filter_small_kernel(const float4* __restrict__ boxes, int K, float min_size,
                                                            uint8_t* __restrict__ keep) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const float4 b = boxes[i];
  keep[i] = (__fsub_rn(b.z, b.x) >= min_size) && (__fsub_rn(b.w, b.y) >= min_size);
}

__global__ void __launch_bounds__(256) decode_kernel(const float4* __restrict__ deltas, const float4* __restrict__ anchors,
                                                      int K, DecodeCfg cfg, float4* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  out[i] = decode_one(__ldg(deltas + i), __ldg(anchors + i), cfg);
}

// a11 — LevelMapper.__call__ (TV:ops/poolers.py:73-84), fp32 throughout:
// floor(lvl0 + log2(sqrt(area)/s0) + eps) clamped to [k_min, k_max], minus k_min.
__global__ void __launch_bounds__(256) level_map_kernel(const float* __restrict__ boxes, int box_stride, int K, int k_min,
                                                         int k_max, float s0, float lvl0, float eps,
                                                         int* __restrict__ levels) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= K) return;
  const float* b = boxes + (size_t)i * box_stride + (box_stride == 5 ? 1 : 0);
  const float area = __fmul_rn(__fsub_rn(b[2], b[0]), __fsub_rn(b[3], b[1]));
  const float s = sqrtf(area);
  float t = floorf(__fadd_rn(__fadd_rn(lvl0, log2f(__fdiv_rn(s, s0))), eps));
  t = t < (float)k_min ? (float)k_min : t;
  t = t > (float)k_max ? (float)k_max : t;
  levels[i] = (t != t) ? 0 : (int)t - k_min;
}

// proposals[keep] gather (src/utils/proposal_utils.py:56-57, src/custom_maskrcnn.py:193-195) + rois.
__global__ void __launch_bounds__(256) gather_kept_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                                                           const int64_t* __restrict__ keep,
                                                           const int* __restrict__ keep_counts, int S, int in_stride,
                                                           int post_n, float4* __restrict__ out_boxes,
                                                           float* __restrict__ out_scores, float* __restrict__ out_rois,
                                                           uint8_t* __restrict__ out_valid) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * post_n) return;
  const int s = i / post_n, j = i - s * post_n;
  const bool live = j < keep_counts[s];
  float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
  float sc = 0.f;
  if (live) {
    const int src = (int)keep[i];
    b = __ldg(boxes + (size_t)s * in_stride + src);
    if (scores) sc = __ldg(scores + (size_t)s * in_stride + src);
  }
  if (out_boxes) out_boxes[i] = b;
  if (out_scores) out_scores[i] = sc;
  if (out_valid) out_valid[i] = live ? 1 : 0;
  if (out_rois) {
    float* r = out_rois + (size_t)i * 5;
    r[0] = live ? (float)s : -1.0f;
    r[1] = b.x; r[2] = b.y; r[3] = b.z; r[4] = b.w;
  }
}

// Detection records (x1,y1,x2,y2,score,label=1) for the all-gather.
__global__ void __launch_bounds__(256) pack_records_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                                                            const int* __restrict__ counts, int S, int stride,
                                                            float* __restrict__ records) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= S * stride) return;
  const int s = i / stride, j = i - s * stride;
  float* r = records + (size_t)i * 6;
  if (j < counts[s]) {
    const float4 b = __ldg(boxes + i);
    r[0] = b.x; r[1] = b.y; r[2] = b.z; r[3] = b.w; r[4] = __ldg(scores + i); r[5] = 1.0f;
  } else {
    r[0] = r[1] = r[2] = r[3] = r[4] = r[5] = 0.f;
  }
}

// Cross-level concatenation for torchvision's RPN (a12; TV:models/detection/rpn.py:258-291, TV:ops/boxes.py:51-120): the
// per-level survivors of lcr_rpn_select_f32 ([B, L, k] padded, counts [B, L]) become one list per image in level order
// (the order filter_proposals sees after its top-n gather and order-preserving filters), with the level id of every box.
// With `trick` the boxes handed to NMS are offset by level * (max coordinate of the image's boxes + 1), exactly as
// _batched_nms_coordinate_trick does (each torch op one fp32 rounding): NMS over the offset boxes then needs no
// category test — and reproduces torchvision's rounding of the shifted coordinates.
__global__ void __launch_bounds__(256) concat_levels_kernel(const float4* __restrict__ boxes, const float* __restrict__ scores,
                                                             const int* __restrict__ counts, int L, int k, int trick,
                                                             float4* __restrict__ cat_boxes, float4* __restrict__ nms_boxes,
                                                             float* __restrict__ cat_scores, int* __restrict__ cat_level,
                                                             int* __restrict__ cat_counts) {
  const int b = blockIdx.x;
  __shared__ int s_off[LCR_MAX_LEVELS + 1];
  __shared__ float s_max[8];
  if (threadIdx.x == 0) {
    int acc = 0;
    for (int l = 0; l < L; ++l) {
      s_off[l] = acc;
      acc += min(max(counts[b * L + l], 0), k);
    }
    s_off[L] = acc;
    cat_counts[b] = acc;
  }
  __syncthreads();
  const int total = s_off[L];
  float shift = 0.f;
  if (trick) {   // boxes.max() over the image's boxes (NaN ignored: fmaxf)
    float m = -INFINITY;
    for (int j = threadIdx.x; j < L * k; j += blockDim.x) {
      const int l = j / k, r = j - l * k;
      if (r < s_off[l + 1] - s_off[l]) {
        const float4 v = __ldg(boxes + ((size_t)b * L + l) * k + r);
        m = fmaxf(m, fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w)));
      }
    }
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) s_max[threadIdx.x >> 5] = m;
    __syncthreads();
    m = s_max[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, s_max[w]);
    shift = __fadd_rn(m, 1.0f);
  }
  for (int j = threadIdx.x; j < L * k; j += blockDim.x) {
    const int l = j / k, r = j - l * k;
    const size_t dst = (size_t)b * L * k + s_off[l] + r;
    if (r < s_off[l + 1] - s_off[l]) {
      const size_t src = ((size_t)b * L + l) * k + r;
      const float4 v = __ldg(boxes + src);
      cat_boxes[dst] = v;
      cat_scores[dst] = __ldg(scores + src);
      cat_level[dst] = l;
      if (nms_boxes) {
        const float o = __fmul_rn((float)l, shift);   // idxs.to(boxes) * (max_coordinate + 1)
        nms_boxes[dst] = trick ? make_float4(__fadd_rn(v.x, o), __fadd_rn(v.y, o), __fadd_rn(v.z, o), __fadd_rn(v.w, o)) : v;
      }
    }
  }
  for (int j = total + threadIdx.x; j < L * k; j += blockDim.x) {   // padding rows
    const size_t dst = (size_t)b * L * k + j;
    cat_boxes[dst] = make_float4(0.f, 0.f, 0.f, 0.f);
    cat_scores[dst] = 0.f;
    cat_level[dst] = -1;
    if (nms_boxes) nms_boxes[dst] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
}

}  // namespace lcr

using namespace lcr;

extern "C" int lcr_anchors_f32(float* out, int h, int w, int stride, const float* base_anchors_host, int A, void* stream) {
  LCR_REQUIRE(out && base_anchors_host && h > 0 && w > 0 && A > 0 && stride > 0, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(A <= LCR_MAX_ANCHORS, LCR_ERR_CAPACITY);
  LCR_REQUIRE((int64_t)h * w * A < (1ll << 31), LCR_ERR_CAPACITY);
  LCR_REQUIRE(aligned_to(out, 16), LCR_ERR_ALIGNMENT);
  BaseAnchors base{};
  for (int i = 0; i < A * 4; ++i) base.v[i] = base_anchors_host[i];
  const int n = h * w * A;
  const int blocks = (n + 255) / 256;
  anchors_kernel<<<blocks, 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(out), h, w, A, (float)stride, base);
  return after_launch();
}

extern "C" int lcr_clip_boxes_f32(float* boxes, int K, float img_h, float img_w, void* stream) {
  LCR_REQUIRE(K >= 0, LCR_ERR_INVALID_ARG);
  if (K == 0) return LCR_OK;
  LCR_REQUIRE(boxes, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(boxes, 16), LCR_ERR_ALIGNMENT);
  clip_kernel<<<(K + 255) / 256, 256, 0, as_stream(stream)>>>(reinterpret_cast<float4*>(boxes), K, img_h, img_w);
  return after_launch();
}

extern "C" int lcr_filter_small_boxes_f32(const float* boxes, int K, float min_size, uint8_t* keep, void* stream) {
  LCR_REQUIRE(K >= 0, LCR_ERR_INVALID_ARG);
  if (K == 0) return LCR_OK;
  LCR_REQUIRE(boxes && keep, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(boxes, 16), LCR_ERR_ALIGNMENT);
  filter_small_kernel<<<(K + 255) / 256, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(boxes), K, min_size, keep);
  return after_launch();
}

extern "C" int lcr_box_decode_f32(const float* deltas, const float* anchors, int K, const float weights_host[4],
                                  float xform_clip, float img_h, float img_w, float* out, void* stream) {
  LCR_REQUIRE(K >= 0 && weights_host, LCR_ERR_INVALID_ARG);
  if (K == 0) return LCR_OK;
  LCR_REQUIRE(deltas && anchors && out, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(deltas, 16) && aligned_to(anchors, 16) && aligned_to(out, 16), LCR_ERR_ALIGNMENT);
  DecodeCfg cfg{weights_host[0], weights_host[1], weights_host[2], weights_host[3], xform_clip, img_h, img_w};
  decode_kernel<<<(K + 255) / 256, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(deltas),
                                                               reinterpret_cast<const float4*>(anchors), K, cfg,
                                                               reinterpret_cast<float4*>(out));
  return after_launch();
}

extern "C" int lcr_level_map_f32(const float* boxes, int box_stride, int K, int k_min, int k_max, float canonical_scale,
                                 int canonical_level, float eps, int* levels, void* stream) {
  LCR_REQUIRE(K >= 0 && (box_stride == 4 || box_stride == 5) && k_max >= k_min, LCR_ERR_INVALID_ARG);
  if (K == 0) return LCR_OK;
  LCR_REQUIRE(boxes && levels, LCR_ERR_INVALID_ARG);
  level_map_kernel<<<(K + 255) / 256, 256, 0, as_stream(stream)>>>(boxes, box_stride, K, k_min, k_max, canonical_scale,
                                                                  (float)canonical_level, eps, levels);
  return after_launch();
}

extern "C" int lcr_gather_kept_f32(const float* boxes, const float* scores, const int64_t* keep, const int* keep_counts,
                                   int S, int in_stride, int post_n, float* out_boxes, float* out_scores, float* out_rois,
                                   uint8_t* out_valid, void* stream) {
  LCR_REQUIRE(S >= 0 && in_stride > 0 && post_n > 0, LCR_ERR_INVALID_ARG);
  if (S == 0) return LCR_OK;
  LCR_REQUIRE(boxes && keep && keep_counts, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(boxes, 16) && (!out_boxes || aligned_to(out_boxes, 16)), LCR_ERR_ALIGNMENT);
  const int n = S * post_n;
  gather_kept_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(boxes), scores, keep,
                                                                    keep_counts, S, in_stride, post_n,
                                                                    reinterpret_cast<float4*>(out_boxes), out_scores, out_rois, out_valid);
  return after_launch();
}

extern "C" int lcr_pack_records_f32(const float* boxes, const float* scores, const int* counts, int S, int stride,
                                    float* records, void* stream) {
  LCR_REQUIRE(S >= 0 && stride > 0, LCR_ERR_INVALID_ARG);
  if (S == 0) return LCR_OK;
  LCR_REQUIRE(boxes && scores && counts && records, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(boxes, 16), LCR_ERR_ALIGNMENT);
  const int n = S * stride;
  pack_records_kernel<<<(n + 255) / 256, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(boxes), scores, counts,
                                                                     S, stride, records);
  return after_launch();
}

extern "C" int lcr_rpn_concat_levels_f32(const float* boxes, const float* scores, const int* counts, int B, int L, int k,
                                         int coordinate_trick, float* cat_boxes, float* nms_boxes, float* cat_scores,
                                         int* cat_level, int* cat_counts, void* stream) {
  LCR_REQUIRE(B >= 0 && L > 0 && k > 0, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(L <= LCR_MAX_LEVELS, LCR_ERR_CAPACITY);
  if (B == 0) return LCR_OK;
  LCR_REQUIRE(boxes && scores && counts && cat_boxes && cat_scores && cat_level && cat_counts, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(!coordinate_trick || nms_boxes, LCR_ERR_INVALID_ARG);
  LCR_REQUIRE(aligned_to(boxes, 16) && aligned_to(cat_boxes, 16) && (!nms_boxes || aligned_to(nms_boxes, 16)), LCR_ERR_ALIGNMENT);
  concat_levels_kernel<<<B, 256, 0, as_stream(stream)>>>(reinterpret_cast<const float4*>(boxes), scores, counts, L, k,
                                                        coordinate_trick ? 1 : 0, reinterpret_cast<float4*>(cat_boxes),
                                                        reinterpret_cast<float4*>(nms_boxes), cat_scores, cat_level, cat_counts);
  return after_launch();
}

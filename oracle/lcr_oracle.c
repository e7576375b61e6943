/*
 * lcr_oracle.c — CPU restatement of the reference's region-pipeline arithmetic.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may build, load or call
 * it, and only as the checker / reported CPU baseline.  The product path (liblcr.so) never links it.
 *
 * Parity pin: the reference has no tests or golden vectors of its own, so this
 * oracle is pinned against outputs of the reference itself, generated in the authoring container by
 * tests/golden/make_golden.py (which imports /root/reference and the installed torchvision 0.26 CPU
 * ops) and committed as tests/golden/ (npz files); tests/test_oracle_golden.py replays them.
 *
 * Where the arithmetic lives in a third-party dependency (torchvision==0.22.1 / torch==2.7.1 pinned in
 * the reference's requirements.txt:2-3; 0.26.0 / 2.11.0 installed here) the published algorithm is
 * restated and the reference's call site is cited.  `TV:` = torchvision python sources.
 *
 * Build: see oracle/Makefile (gcc -O2 -ffp-contract=off [-fopenmp]).  -ffp-contract=off matters:
 * every fused multiply-add below is written explicitly with fmaf().
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

/* Launchers such as torchrun export OMP_NUM_THREADS=1 to every rank; the CPU baseline leg of bench.py asks for the
 * host's cores explicitly instead of inheriting that. */
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---------------------------------------------------------------------------------------------
 * Anchors — src/components/anchor_generator.py:13-37.
 * base[a] (already rounded to fp32 by the caller, as torch.tensor(..., float32) does at :27) is
 * added to fp32 shifts (x*stride, y*stride, x*stride, y*stride) (:29-34); order (y, x, a).
 * ------------------------------------------------------------------------------------------- */
void orc_anchors(float* out, int h, int w, int stride, const float* base, int A) {
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const float sx = (float)x * (float)stride, sy = (float)y * (float)stride;
      for (int a = 0; a < A; ++a) {
        float* o = out + ((size_t)(y * w + x) * A + a) * 4;
        o[0] = sx + base[a * 4 + 0];
        o[1] = sy + base[a * 4 + 1];
        o[2] = sx + base[a * 4 + 2];
        o[3] = sy + base[a * 4 + 3];
      }
    }
}

/* clip_boxes_to_image — src/utils/box_utils.py:32-37 (x to [0,w], y to [0,h], in place). */
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
void orc_clip_boxes(float* boxes, int K, float img_h, float img_w) {
  for (int i = 0; i < K; ++i) {
    float* b = boxes + (size_t)i * 4;
    b[0] = clampf(b[0], 0.f, img_w);
    b[2] = clampf(b[2], 0.f, img_w);
    b[1] = clampf(b[1], 0.f, img_h);
    b[3] = clampf(b[3], 0.f, img_h);
  }
}

/* filter_small_boxes — src/utils/box_utils.py:39-44. */
void orc_filter_small_boxes(const float* boxes, int K, float min_size, uint8_t* keep) {
  for (int i = 0; i < K; ++i) {
    const float* b = boxes + (size_t)i * 4;
    keep[i] = ((b[2] - b[0]) >= min_size) && ((b[3] - b[1]) >= min_size);
  }
}

/* BoxCoder.decode_single — TV:models/detection/_utils.py:183-224 (the inverse of encode_boxes,
 * src/utils/box_utils.py:4-28).  Every product is a separate fp32 rounding, as in the torch op chain. */
void orc_box_decode(const float* deltas, const float* anchors, int K, const float* wts, float xform_clip,
                    float img_h, float img_w, float* out) {
  for (int i = 0; i < K; ++i) {
    const float* a = anchors + (size_t)i * 4;
    const float* d = deltas + (size_t)i * 4;
    const float w = a[2] - a[0], h = a[3] - a[1];
    const float cx = a[0] + 0.5f * w, cy = a[1] + 0.5f * h;
    float dx = d[0] / wts[0], dy = d[1] / wts[1], dw = d[2] / wts[2], dh = d[3] / wts[3];
    if (dw > xform_clip) dw = xform_clip;
    if (dh > xform_clip) dh = xform_clip;
    const float pcx = dx * w + cx, pcy = dy * h + cy;
    const float pw = expf(dw) * w, ph = expf(dh) * h;
    const float hw = 0.5f * pw, hh = 0.5f * ph;
    float* o = out + (size_t)i * 4;
    o[0] = pcx - hw; o[1] = pcy - hh; o[2] = pcx + hw; o[3] = pcy + hh;
    if (img_h > 0.f) {
      o[0] = clampf(o[0], 0.f, img_w); o[2] = clampf(o[2], 0.f, img_w);
      o[1] = clampf(o[1], 0.f, img_h); o[3] = clampf(o[3], 0.f, img_h);
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * RPN proposal selection for ONE (image, level) segment —
 * src/utils/proposal_utils.py:16-29 (training) / :38-52 (inference), up to but excluding NMS.
 * torchvision variant (topk_on_sigmoid = 0, deltas given): TV:models/detection/rpn.py:231-297.
 *
 * obj: [A, h, w] logits of this image/level.  Flat index i = (y*w + x)*A + a.
 * Tie rule (ours; torch.topk leaves it unspecified): (key desc, flat index asc), NaN highest.
 * Returns the number of surviving proposals; boxes/scores/index have capacity k.
 * ------------------------------------------------------------------------------------------- */
static inline float sigmoidf_(float x) { return 1.0f / (1.0f + expf(-x)); }

/* Order-preserving map float -> uint32 (ascending), -0 == +0, every NaN on top. */
static inline uint32_t order_key(float v) {
  if (v != v) return 0xFFFFFFFFu;
  v = v + 0.0f;
  uint32_t u;
  memcpy(&u, &v, 4);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

typedef struct { uint32_t key; uint32_t idx; } KeyIdx;
static int cmp_keyidx_desc(const void* pa, const void* pb) {
  const KeyIdx* a = (const KeyIdx*)pa; const KeyIdx* b = (const KeyIdx*)pb;
  if (a->key != b->key) return a->key > b->key ? -1 : 1;
  if (a->idx != b->idx) return a->idx < b->idx ? -1 : 1;
  return 0;
}

int orc_rpn_select_segment(const float* obj, const float* deltas, const float* anchors,
                           const float* base_anchors, int A, int h, int w, int stride, int k,
                           float score_thresh, int score_strict, float min_size, int img_h, int img_w,
                           int topk_on_sigmoid, const float* decode_weights, float xform_clip,
                           float* boxes, float* scores, int64_t* index) {
  const int hw = h * w;
  const int n = A * hw;
  KeyIdx* e = (KeyIdx*)malloc(sizeof(KeyIdx) * (size_t)n);
  for (int a = 0; a < A; ++a)
    for (int p = 0; p < hw; ++p) {
      const float logit = obj[(size_t)a * hw + p];
      const int i = p * A + a;
      e[i].key = order_key(topk_on_sigmoid ? sigmoidf_(logit) : logit);
      e[i].idx = (uint32_t)i;
    }
  qsort(e, (size_t)n, sizeof(KeyIdx), cmp_keyidx_desc);
  const int kk = k < n ? k : n;
  int cnt = 0;
  for (int j = 0; j < kk; ++j) {
    const int i = (int)e[j].idx;
    const int a = i % A, p = i / A, y = p / w, x = p % w;
    const float score = sigmoidf_(obj[(size_t)a * hw + p]);
    /* `score > thr` with NaN: comparisons with NaN are false, so a NaN score is dropped. */
    const int pass = score_strict ? (score > score_thresh) : (score >= score_thresh);
    if (!pass) continue;
    float b[4];
    if (anchors) {
      memcpy(b, anchors + (size_t)i * 4, 16);
    } else {
      const float sx = (float)x * (float)stride, sy = (float)y * (float)stride;
      b[0] = sx + base_anchors[a * 4 + 0]; b[1] = sy + base_anchors[a * 4 + 1];
      b[2] = sx + base_anchors[a * 4 + 2]; b[3] = sy + base_anchors[a * 4 + 3];
    }
    if (deltas) {
      float d[4], o[4];
      for (int c = 0; c < 4; ++c) d[c] = deltas[((size_t)(a * 4 + c)) * hw + p];
      orc_box_decode(d, b, 1, decode_weights, xform_clip, 0.f, 0.f, o);
      memcpy(b, o, 16);
    }
    orc_clip_boxes(b, 1, (float)img_h, (float)img_w);
    if (!(((b[2] - b[0]) >= min_size) && ((b[3] - b[1]) >= min_size))) continue;
    memcpy(boxes + (size_t)cnt * 4, b, 16);
    scores[cnt] = score;
    index[cnt] = i;
    ++cnt;
  }
  free(e);
  return cnt;
}

/* Batched driver (OpenMP over segments) with the liblcr layout: single level, [B, A, h, w]. */
void orc_rpn_select_batch(const float* obj, const float* anchors, const float* base_anchors, int B, int A,
                          int h, int w, int stride, int k, float score_thresh, int score_strict,
                          float min_size, int img_h, int img_w, int topk_on_sigmoid, float* boxes,
                          float* scores, int64_t* index, int* counts) {
  const float wts[4] = {1.f, 1.f, 1.f, 1.f};
#pragma omp parallel for schedule(dynamic, 1)
  for (int b = 0; b < B; ++b)
    counts[b] = orc_rpn_select_segment(obj + (size_t)b * A * h * w, NULL, anchors, base_anchors, A, h, w,
                                       stride, k, score_thresh, score_strict, min_size, img_h, img_w,
                                       topk_on_sigmoid, wts, 4.135166556742356f,
                                       boxes + (size_t)b * k * 4, scores + (size_t)b * k,
                                       index + (size_t)b * k);
}

/* ---------------------------------------------------------------------------------------------
 * Greedy NMS — torchvision.ops.nms (TV:ops/boxes.py:20-48) as called at
 * src/utils/proposal_utils.py:55 and src/custom_maskrcnn.py:192; algorithm per SURVEY.md App. B.3:
 * stable descending sort by score (ties -> lower index, NaN first); box j suppressed by an earlier kept
 * box i iff (double)(inter / (area_i + area_j - inter)) > thr, all IoU arithmetic in fp32.
 * scores == NULL: input order is the candidate order.  category (nullable): only same-category boxes
 * interact (batched_nms per-class branch, TV:ops/boxes.py:107-120).  use_score_thresh: drop
 * score <= score_thresh first (src/custom_maskrcnn.py:185-188).  Returns #kept written (<= post_n).
 * ------------------------------------------------------------------------------------------- */
static int cmp_score_desc_stable(const void* pa, const void* pb) { return cmp_keyidx_desc(pa, pb); }

int orc_nms(const float* boxes, const float* scores, const int* category, int n, double iou_thr,
            float score_thresh, int use_score_thresh, int post_n, int64_t* keep) {
  if (n <= 0) return 0;
  KeyIdx* order = (KeyIdx*)malloc(sizeof(KeyIdx) * (size_t)n);
  int m = 0;
  for (int i = 0; i < n; ++i) {
    if (scores && use_score_thresh && !(scores[i] > score_thresh)) continue;
    order[m].key = scores ? order_key(scores[i]) : 0u;
    order[m].idx = (uint32_t)i;
    ++m;
  }
  if (scores) qsort(order, (size_t)m, sizeof(KeyIdx), cmp_score_desc_stable);
  uint8_t* dead = (uint8_t*)calloc((size_t)(m > 0 ? m : 1), 1);
  float* area = (float*)malloc(sizeof(float) * (size_t)(m > 0 ? m : 1));
  for (int r = 0; r < m; ++r) {
    const float* b = boxes + (size_t)order[r].idx * 4;
    area[r] = (b[2] - b[0]) * (b[3] - b[1]);
  }
  int cnt = 0;
  for (int r = 0; r < m && cnt < post_n; ++r) {
    if (dead[r]) continue;
    const int i = (int)order[r].idx;
    keep[cnt++] = i;
    const float* bi = boxes + (size_t)i * 4;
    for (int q = r + 1; q < m; ++q) {
      if (dead[q]) continue;
      const int j = (int)order[q].idx;
      if (category && category[i] != category[j]) continue;
      const float* bj = boxes + (size_t)j * 4;
      const float xx1 = bi[0] > bj[0] ? bi[0] : bj[0];
      const float yy1 = bi[1] > bj[1] ? bi[1] : bj[1];
      const float xx2 = bi[2] < bj[2] ? bi[2] : bj[2];
      const float yy2 = bi[3] < bj[3] ? bi[3] : bj[3];
      float iw = xx2 - xx1; if (!(iw > 0.f)) iw = 0.f;
      float ih = yy2 - yy1; if (!(ih > 0.f)) ih = 0.f;
      const float inter = iw * ih;
      const float iou = inter / (area[r] + area[q] - inter);
      if ((double)iou > iou_thr) dead[q] = 1;
    }
  }
  free(order); free(dead); free(area);
  return cnt;
}

void orc_nms_batch(const float* boxes, const float* scores, const int* counts, int S, int stride,
                   double iou_thr, float score_thresh, int use_score_thresh, int post_n, int64_t* keep,
                   int* keep_counts) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int s = 0; s < S; ++s)
    keep_counts[s] = orc_nms(boxes + (size_t)s * stride * 4, scores ? scores + (size_t)s * stride : NULL,
                             NULL, counts ? counts[s] : stride, iou_thr, score_thresh, use_score_thresh,
                             post_n, keep + (size_t)s * post_n);
}

/* LevelMapper.__call__ — TV:ops/poolers.py:73-84 (fp32 throughout). */
void orc_level_map(const float* boxes, int box_stride, int K, int k_min, int k_max, float s0, int lvl0,
                   float eps, int* levels) {
  const int off = box_stride == 5 ? 1 : 0;
  for (int i = 0; i < K; ++i) {
    const float* b = boxes + (size_t)i * box_stride + off;
    const float area = (b[2] - b[0]) * (b[3] - b[1]);
    const float s = sqrtf(area);
    float t = floorf((float)lvl0 + log2f(s / s0) + eps);
    if (t < (float)k_min) t = (float)k_min;
    if (t > (float)k_max) t = (float)k_max;
    /* NaN area (never produced by clipped boxes) would convert as torch's .to(int64) does: UB; pin to k_min. */
    levels[i] = (t != t) ? 0 : (int)t - k_min;
  }
}

/* ---------------------------------------------------------------------------------------------
 * RoIAlign forward — torchvision::roi_align as called by RoIAlign((7,7), 0.25, 2) at
 * src/custom_maskrcnn.py:48-50,:120,:177; kernel math per SURVEY.md App. B.1 (bit-identical to the
 * compiled CPU op there), including the compiled-kernel-only rule that samples with
 * y < -1 || y > H || x < -1 || x > W contribute 0 but stay in `count`.
 * feat: logical [N, C, H, W] with element strides (sn, sc, sh, sw); rois [K,5]; out [K,C,PH,PW].
 * rois with batch index < 0 are padding -> zeros.
 * ------------------------------------------------------------------------------------------- */
typedef struct { int yl, yh, xl, xh; float w1, w2, w3, w4; int valid; } Tap;

/* `aligned` is a flag word: bit 0 = torchvision's aligned, bit 1 = the sample coordinates are rounded as torchvision's CUDA
 * kernel rounds them (nvcc contracts `roi * spatial_scale - offset` and `roi_start + ph * bin_size` into FMAs) instead of as
 * its CPU kernel does (every operation rounded).  The two ops of the same torchvision differ by an ulp of the coordinate,
 * i.e. by up to ~2e-5 of the output range on white-noise features (measured on the B200: tools/roi_coord_rounding_exp.py);
 * the golden vectors are CPU-generated (bit 1 clear), the reference on a GPU runs the CUDA op (bit 1 set). */
static inline float sample_coord(float start, int bin_index, float bin, int i, int grid, int cuda_coords) {
  const float head = cuda_coords ? fmaf((float)bin_index, bin, start) : start + (float)bin_index * bin;
  return head + ((float)i + 0.5f) * bin / (float)grid;
}

static void roi_geometry(const float* roi, float scale, int aligned_flags, int PH, int PW, int sr, float* sw,
                         float* sh, float* bw, float* bh, int* gh, int* gw) {
  const int aligned = aligned_flags & 1, cuda_coords = (aligned_flags >> 1) & 1;
  const float off = aligned ? 0.5f : 0.0f;
  float ew, eh;
  if (cuda_coords) {
    *sw = fmaf(roi[1], scale, -off); *sh = fmaf(roi[2], scale, -off);
    ew = fmaf(roi[3], scale, -off); eh = fmaf(roi[4], scale, -off);
  } else {
    *sw = roi[1] * scale - off; *sh = roi[2] * scale - off;
    ew = roi[3] * scale - off; eh = roi[4] * scale - off;
  }
  float rw = ew - *sw, rh = eh - *sh;
  if (!aligned) { rw = rw > 1.f ? rw : 1.f; rh = rh > 1.f ? rh : 1.f; }
  *bh = rh / (float)PH; *bw = rw / (float)PW;
  *gh = sr > 0 ? sr : (int)ceilf(rh / (float)PH);
  *gw = sr > 0 ? sr : (int)ceilf(rw / (float)PW);
}

static inline Tap make_tap(float y, float x, int H, int W) {
  Tap t; memset(&t, 0, sizeof(t));
  if (y < -1.0f || y > (float)H || x < -1.0f || x > (float)W) return t;
  if (y <= 0.f) y = 0.f;
  if (x <= 0.f) x = 0.f;
  int yl = (int)y, xl = (int)x, yh, xh;
  if (yl >= H - 1) { yh = yl = H - 1; y = (float)yl; } else yh = yl + 1;
  if (xl >= W - 1) { xh = xl = W - 1; x = (float)xl; } else xh = xl + 1;
  const float ly = y - (float)yl, lx = x - (float)xl, hy = 1.f - ly, hx = 1.f - lx;
  t.yl = yl; t.yh = yh; t.xl = xl; t.xh = xh;
  t.w1 = hy * hx; t.w2 = hy * lx; t.w3 = ly * hx; t.w4 = ly * lx; t.valid = 1;
  return t;
}

void orc_roi_align_fwd(const float* feat, int N, int C, int H, int W, int64_t sn, int64_t sc, int64_t sh_,
                       int64_t sw_, const float* rois, int K, int PH, int PW, float scale, int sr,
                       int aligned, float* out) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int k = 0; k < K; ++k) {
    const float* roi = rois + (size_t)k * 5;
    float* o = out + (size_t)k * C * PH * PW;
    const int b = (int)roi[0];
    if (roi[0] < 0.f || b >= N) { memset(o, 0, sizeof(float) * (size_t)C * PH * PW); continue; }
    float sw, sh, bw, bh; int gh, gw;
    roi_geometry(roi, scale, aligned, PH, PW, sr, &sw, &sh, &bw, &bh, &gh, &gw);
    const float count = (float)((gh * gw) > 1 ? gh * gw : 1);
    const float* f0 = feat + (size_t)b * sn;
    for (int ph = 0; ph < PH; ++ph)
      for (int pw = 0; pw < PW; ++pw) {
        for (int c = 0; c < C; ++c) o[((size_t)c * PH + ph) * PW + pw] = 0.f;
        for (int iy = 0; iy < gh; ++iy) {
          const float y = sample_coord(sh, ph, bh, iy, gh, (aligned >> 1) & 1);
          for (int ix = 0; ix < gw; ++ix) {
            const float x = sample_coord(sw, pw, bw, ix, gw, (aligned >> 1) & 1);
            const Tap t = make_tap(y, x, H, W);
            if (!t.valid) continue;
            for (int c = 0; c < C; ++c) {
              const float* f = f0 + (size_t)c * sc;
              /* left-to-right sum of four separately rounded products (App. B.1) */
              float v = t.w1 * f[t.yl * sh_ + t.xl * sw_];
              v = v + t.w2 * f[t.yl * sh_ + t.xh * sw_];
              v = v + t.w3 * f[t.yh * sh_ + t.xl * sw_];
              v = v + t.w4 * f[t.yh * sh_ + t.xh * sw_];
              o[((size_t)c * PH + ph) * PW + pw] += v;
            }
          }
        }
        for (int c = 0; c < C; ++c) o[((size_t)c * PH + ph) * PW + pw] /= count;
      }
  }
}

/* RoIAlign backward — torchvision::_roi_align_backward (reached by autograd from
 * src/train_custom.py:44); App. B.2: each sample adds g*w/count to its four neighbours,
 * sequential order k -> ph -> pw -> iy -> ix.  grad_in must be zeroed by the caller. */
void orc_roi_align_bwd(const float* grad_out, int N, int C, int H, int W, int64_t sn, int64_t sc,
                       int64_t sh_, int64_t sw_, const float* rois, int K, int PH, int PW, float scale,
                       int sr, int aligned, float* grad_in) {
  for (int k = 0; k < K; ++k) {
    const float* roi = rois + (size_t)k * 5;
    const int b = (int)roi[0];
    if (roi[0] < 0.f || b >= N) continue;
    float sw, sh, bw, bh; int gh, gw;
    roi_geometry(roi, scale, aligned, PH, PW, sr, &sw, &sh, &bw, &bh, &gh, &gw);
    const float count = (float)((gh * gw) > 1 ? gh * gw : 1);
    float* g0 = grad_in + (size_t)b * sn;
    const float* go = grad_out + (size_t)k * C * PH * PW;
    for (int ph = 0; ph < PH; ++ph)
      for (int pw = 0; pw < PW; ++pw)
        for (int iy = 0; iy < gh; ++iy) {
          const float y = sample_coord(sh, ph, bh, iy, gh, (aligned >> 1) & 1);
          for (int ix = 0; ix < gw; ++ix) {
            const float x = sample_coord(sw, pw, bw, ix, gw, (aligned >> 1) & 1);
            const Tap t = make_tap(y, x, H, W);
            if (!t.valid) continue;
            for (int c = 0; c < C; ++c) {
              const float g = go[((size_t)c * PH + ph) * PW + pw];
              float* gi = g0 + (size_t)c * sc;
              gi[t.yl * sh_ + t.xl * sw_] += g * t.w1 / count;
              gi[t.yl * sh_ + t.xh * sw_] += g * t.w2 / count;
              gi[t.yh * sh_ + t.xl * sw_] += g * t.w3 / count;
              gi[t.yh * sh_ + t.xh * sw_] += g * t.w4 / count;
            }
          }
        }
  }
}

/* ---------------------------------------------------------------------------------------------
 * Mask paste — CustomMaskRCNN._generate_masks (src/custom_maskrcnn.py:276-295) and its twin
 * paste_masks_in_image (src/utils/mask_utils.py:149-171): int-truncated box, clamp to the image,
 * bilinear (align_corners=False) resize of the MxM probability map to the box, `> thr`, {0,on}.
 * Bilinear per SURVEY.md App. B.4 with the FMA placement found for ATen's CPU kernel:
 *   src = max(fma(in/out, dst+0.5, -0.5), 0); i0 = (int)src; i1 = i0 + (i0 < in-1); l1 = src-i0; l0 = 1-l1
 *   row = fma(a, wx0, b*wx1);  val = fma(row0, wy0, row1*wy1)
 * (only the thresholded result is claimed bit-exact).
 * valid (nullable): frames with valid[i]==0 are left untouched.
 * Reference quirk not reproduced: `.squeeze()` at custom_maskrcnn.py:290 makes a (bh>1, bw==1) box
 * raise a broadcast error; such boxes cannot reach it through the model (min_box_size filter).
 * ------------------------------------------------------------------------------------------- */
static inline void src_index(float scale, int dst, int in_size, int* i0, int* i1, float* l0, float* l1) {
  float src = fmaf(scale, (float)dst + 0.5f, -0.5f);
  if (src < 0.f) src = 0.f;
  int a = (int)src;
  if (a > in_size - 1) a = in_size - 1;
  *i0 = a;
  *i1 = a + ((a < in_size - 1) ? 1 : 0);
  *l1 = src - (float)a;
  *l0 = 1.0f - *l1;
}

void orc_paste_masks(const float* probs, const float* boxes, const uint8_t* valid, int N, int M, int H,
                     int W, float thr, uint8_t on_value, uint8_t* out) {
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < N; ++i) {
    if (valid && !valid[i]) continue;
    uint8_t* frame = out + (size_t)i * H * W;
    memset(frame, 0, (size_t)H * W);
    const float* b = boxes + (size_t)i * 4;
    int x1 = (int)b[0], y1 = (int)b[1], x2 = (int)b[2], y2 = (int)b[3]; /* Tensor.int(): truncation */
    if (x1 < 0) x1 = 0;
    if (y1 < 0) y1 = 0;
    if (x2 > W) x2 = W;
    if (y2 > H) y2 = H;
    if (!(x2 > x1 && y2 > y1)) continue;
    const int bh = y2 - y1, bw = x2 - x1;
    const float sh = (float)M / (float)bh, sw = (float)M / (float)bw;
    const float* p = probs + (size_t)i * M * M;
    for (int y = 0; y < bh; ++y) {
      int h0, h1; float wy0, wy1;
      src_index(sh, y, M, &h0, &h1, &wy0, &wy1);
      for (int x = 0; x < bw; ++x) {
        int w0, w1; float wx0, wx1;
        src_index(sw, x, M, &w0, &w1, &wx0, &wx1);
        const float r0 = fmaf(p[h0 * M + w0], wx0, p[h0 * M + w1] * wx1);
        const float r1 = fmaf(p[h1 * M + w0], wx0, p[h1 * M + w1] * wx1);
        const float v = fmaf(r0, wy0, r1 * wy1);
        frame[(size_t)(y1 + y) * W + (x1 + x)] = (v > thr) ? on_value : 0;
      }
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * torchvision's paste (a13, P2 variant) — paste_masks_in_image / expand_masks / expand_boxes / paste_mask_in_image,
 * TV:models/detection/roi_heads.py:405-501.  scale = float(M + 2 pad) / M is a Python float multiplied into a float32
 * tensor (rounded to fp32); boxes_exp.to(int64) truncates; w = max(int(x2 - x1 + 1), 1); the padded map is resized with
 * ATen's bilinear (same src_index / FMA placement as above); the frame receives
 * mask[(y_0 - y1):(y_1 - y1), (x_0 - x1):(x_1 - x1)] with x_0 = max(x1, 0), x_1 = min(x2 + 1, W).
 * Boxes that lie entirely outside the frame leave it zero (torchvision's negative slice bounds would wrap around; its
 * callers clip boxes to the image first).  out [N, H, W] float32.
 * ------------------------------------------------------------------------------------------- */
void orc_paste_masks_tv(const float* probs, const float* boxes, int N, int M, int H, int W, int pad, float* out) {
  const int Mp = M + 2 * pad;
  const float scale = (float)((double)Mp / (double)M);
#pragma omp parallel for schedule(dynamic, 1)
  for (int i = 0; i < N; ++i) {
    float* frame = out + (size_t)i * H * W;
    memset(frame, 0, sizeof(float) * (size_t)H * W);
    float* pm = (float*)calloc((size_t)Mp * Mp, sizeof(float));
    for (int y = 0; y < M; ++y)
      for (int x = 0; x < M; ++x) pm[(y + pad) * Mp + (x + pad)] = probs[(size_t)i * M * M + y * M + x];
    const float* b = boxes + (size_t)i * 4;
    const float wh = ((b[2] - b[0]) * 0.5f) * scale, hh = ((b[3] - b[1]) * 0.5f) * scale;
    const float xc = (b[2] + b[0]) * 0.5f, yc = (b[3] + b[1]) * 0.5f;
    const long long x1 = (long long)(xc - wh), x2 = (long long)(xc + wh), y1 = (long long)(yc - hh), y2 = (long long)(yc + hh);
    long long w = x2 - x1 + 1, h = y2 - y1 + 1;
    if (w < 1) w = 1;
    if (h < 1) h = 1;
    const float sh = (float)Mp / (float)h, sw = (float)Mp / (float)w;
    const long long xa = x1 > 0 ? x1 : 0, xb = (x2 + 1 < W) ? x2 + 1 : W;   /* x_1 = min(box[2] + 1, im_w): empty for x2 < x1 */
    const long long ya = y1 > 0 ? y1 : 0, yb = (y2 + 1 < H) ? y2 + 1 : H;
    for (long long y = ya; y < yb; ++y) {
      int h0, h1; float wy0, wy1;
      src_index(sh, (int)(y - y1), Mp, &h0, &h1, &wy0, &wy1);
      for (long long x = xa; x < xb; ++x) {
        int w0, w1; float wx0, wx1;
        src_index(sw, (int)(x - x1), Mp, &w0, &w1, &wx0, &wx1);
        const float r0 = fmaf(pm[h0 * Mp + w0], wx0, pm[h0 * Mp + w1] * wx1);
        const float r1 = fmaf(pm[h1 * Mp + w0], wx0, pm[h1 * Mp + w1] * wx1);
        frame[(size_t)y * W + x] = fmaf(r0, wy0, r1 * wy1);
      }
    }
    free(pm);
  }
}

/* Detection records (x1,y1,x2,y2,score,label=1) padded with zeros — SURVEY.md §8(e);
 * labels = ones per src/custom_maskrcnn.py:204. */
void orc_pack_records(const float* boxes, const float* scores, const int* counts, int S, int stride,
                      float* records) {
  for (int s = 0; s < S; ++s)
    for (int j = 0; j < stride; ++j) {
      float* r = records + ((size_t)s * stride + j) * 6;
      if (j < counts[s]) {
        memcpy(r, boxes + ((size_t)s * stride + j) * 4, 16);
        r[4] = scores[(size_t)s * stride + j];
        r[5] = 1.0f;
      } else {
        memset(r, 0, 24);
      }
    }
}

/* ---------------------------------------------------------------------------------------------
 * torchvision.ops.box_iou (TV:ops/boxes.py:308-370) + `.max(dim=1)`
 * (src/components/rpn.py:72-73, src/custom_maskrcnn.py:221-222,249-250, src/utils/mask_utils.py:93-94).
 * iou may be NULL (only the row max / argmax are wanted) and so may max_iou/argmax.
 * ------------------------------------------------------------------------------------------- */
void orc_box_iou(const float* boxes, int N, const float* gt, int G, float* iou, float* max_iou, int64_t* argmax) {
#pragma omp parallel for schedule(static)
  for (int i = 0; i < N; ++i) {
    const float* a = boxes + (size_t)i * 4;
    const float area_a = (a[2] - a[0]) * (a[3] - a[1]);
    float best = 0.f;
    int best_j = 0, have = 0, best_nan = 0;
    for (int j = 0; j < G; ++j) {
      const float* b = gt + (size_t)j * 4;
      const float area_b = (b[2] - b[0]) * (b[3] - b[1]);
      const float left = a[0] > b[0] ? a[0] : b[0], right = a[2] < b[2] ? a[2] : b[2];
      const float top = a[1] > b[1] ? a[1] : b[1], bottom = a[3] < b[3] ? a[3] : b[3];
      float w = right - left, h = bottom - top;
      if (!(w > 0.f)) w = 0.f; /* clamp(min=0) */
      if (!(h > 0.f)) h = 0.f;
      const float inter = w * h;
      const float v = inter / ((area_a + area_b) - inter);
      if (iou) iou[(size_t)i * G + j] = v;
      if (!best_nan && (!have || v > best || v != v)) { /* torch.max: first maximum; NaN propagates */
        best = v;
        best_j = j;
        best_nan = (v != v);
        have = 1;
      }
    }
    if (max_iou) max_iou[i] = best;
    if (argmax) argmax[i] = best_j;
  }
}

/* extract_mask_target (src/utils/mask_utils.py:6-46) for K (box, mask index) pairs. */
void orc_mask_targets(const uint8_t* gt_masks, int G, int H, int W, const float* boxes, const int64_t* gt_index, int K,
                      int M, float* out) {
#pragma omp parallel for schedule(dynamic, 4)
  for (int k = 0; k < K; ++k) {
    float* o = out + (size_t)k * M * M;
    const int64_t gi = gt_index ? gt_index[k] : k;
    if (gi < 0 || gi >= G) {
      memset(o, 0, sizeof(float) * M * M);
      continue;
    }
    const float* b = boxes + (size_t)k * 4;
    int x1 = (int)b[0], y1 = (int)b[1], x2 = (int)b[2], y2 = (int)b[3]; /* box.int(): truncation */
    x1 = x1 < W - 1 ? x1 : W - 1; if (x1 < 0) x1 = 0;                    /* max(0, min(x1, w - 1)) */
    y1 = y1 < H - 1 ? y1 : H - 1; if (y1 < 0) y1 = 0;
    x2 = x2 < W ? x2 : W; if (x2 < x1 + 1) x2 = x1 + 1;                  /* max(x1 + 1, min(x2, w)) */
    y2 = y2 < H ? y2 : H; if (y2 < y1 + 1) y2 = y1 + 1;
    const int ch = y2 - y1, cw = x2 - x1;
    const float sh = (float)ch / (float)M, sw = (float)cw / (float)M;
    const uint8_t* m = gt_masks + (size_t)gi * H * W + (size_t)y1 * W + x1;
    for (int oy = 0; oy < M; ++oy) {
      int h0, h1; float wy0, wy1;
      src_index(sh, oy, ch, &h0, &h1, &wy0, &wy1);
      for (int ox = 0; ox < M; ++ox) {
        int w0, w1; float wx0, wx1;
        src_index(sw, ox, cw, &w0, &w1, &wx0, &wx1);
        const float r0 = fmaf((float)m[(size_t)h0 * W + w0], wx0, (float)m[(size_t)h0 * W + w1] * wx1);
        const float r1 = fmaf((float)m[(size_t)h1 * W + w0], wx0, (float)m[(size_t)h1 * W + w1] * wx1);
        o[oy * M + ox] = fmaf(r0, wy0, r1 * wy1);
      }
    }
  }
}

/* tail of CustomMaskHead.forward (src/components/mask_head.py:52-58: bilinear
 * resize of the logits, align_corners=False) + sigmoid of class `cls` (src/custom_maskrcnn.py:273-274). */
void orc_mask_tail(const float* logits, int K, int num_classes, int cls, int m, int M, float* probs) {
  const float scale = (float)m / (float)M;
#pragma omp parallel for schedule(static)
  for (int k = 0; k < K; ++k) {
    const float* src = logits + ((size_t)k * num_classes + cls) * m * m;
    float* o = probs + (size_t)k * M * M;
    for (int oy = 0; oy < M; ++oy)
      for (int ox = 0; ox < M; ++ox) {
        float v;
        if (m == M) {
          v = src[oy * m + ox];
        } else {
          int h0, h1, w0, w1; float wy0, wy1, wx0, wx1;
          src_index(scale, oy, m, &h0, &h1, &wy0, &wy1);
          src_index(scale, ox, m, &w0, &w1, &wx0, &wx1);
          const float r0 = fmaf(src[h0 * m + w0], wx0, src[h0 * m + w1] * wx1);
          const float r1 = fmaf(src[h1 * m + w0], wx0, src[h1 * m + w1] * wx1);
          v = fmaf(r0, wy0, r1 * wy1);
        }
        o[oy * M + ox] = sigmoidf_(v);
      }
  }
}

/* the pixel counts behind calculate_mask_area_in_region (src/visualize.py:106-130). */
void orc_mask_region_counts(const uint8_t* masks, int N, int H, int W, const int* rects, const int* rect_offsets,
                            int threshold, int* total, int* in_region) {
  for (int i = 0; i < N; ++i) {
    const uint8_t* m = masks + (size_t)i * H * W;
    int t = 0;
    for (int p = 0; p < H * W; ++p) t += ((int)m[p] > threshold);
    total[i] = t;
    for (int r = rect_offsets[i]; r < rect_offsets[i + 1]; ++r) {
      const int* q = rects + (size_t)r * 4;
      int c = 0;
      for (int y = q[1]; y < q[3]; ++y)
        for (int x = q[0]; x < q[2]; ++x) c += ((int)m[(size_t)y * W + x] > threshold);
      in_region[r] = c;
    }
  }
}

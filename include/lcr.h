/*
 * lcr.h — C-ABI of the B200-native LIVECell region pipeline (liblcr.so, sm_100a only).
 *
 * This is the drop-in boundary of the repo.  The reference
 * (jakubradziejewski/livecell-instance-segmentation) has no FFI of its own: its boundary is a set of
 * Python callables plus three torchvision dispatcher ops.  Every entry point below cites the reference
 * interface (file:line under the reference checkout, or `TV:` = torchvision 0.22/0.26 python sources)
 * whose arithmetic it replaces.  INTEGRATION.md shows the ctypes stub a reference maintainer would add.
 *
 * Conventions
 *   - `extern "C"`, plain pointers and sizes, no torch types.  All tensor pointers are DEVICE pointers
 *     unless a parameter is documented as HOST.  All tensors are dense row-major unless strides are given.
 *   - The library never allocates, frees or retains memory: outputs and workspace are caller-owned
 *     (query sizes with lcr_*_workspace_bytes).  Workspace must be 256-byte aligned.
 *   - All work is enqueued on `stream` (a cudaStream_t passed as void*; NULL = legacy default stream).
 *     No entry point synchronises the device or the stream; all are CUDA-graph capturable.
 *   - Variable-length results are written into fixed-capacity padded buffers plus device-side int
 *     counts (no host sync inside the library).
 *   - Return value: LCR_OK (0) or a negative LcrStatus.  Argument errors are detected on the host
 *     before any launch; launch failures are reported as LCR_ERR_CUDA (query lcr_last_cuda_error()).
 *   - There is NO CPU fallback anywhere behind this ABI.
 */
#ifndef LCR_H_
#define LCR_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LCR_VERSION 100          /* 0.1.0 */
#define LCR_MAX_ANCHORS 16       /* anchors per location (reference: 9, src/components/anchor_generator.py:11) */
#define LCR_MAX_LEVELS 8         /* FPN levels per call (reference: 4, src/custom_maskrcnn.py:45) */
#define LCR_MAX_TOPK 8192        /* pre-NMS top-k capacity per (image, level) segment */
#define LCR_MAX_NMS_BOXES 32768  /* boxes per NMS segment */

typedef enum LcrStatus {
  LCR_OK = 0,
  LCR_ERR_INVALID_ARG = -1,   /* null pointer, non-positive size, unsupported parameter combination */
  LCR_ERR_CAPACITY = -2,      /* a size exceeds a compiled-in capacity (LCR_MAX_*) */
  LCR_ERR_WORKSPACE = -3,     /* workspace missing, misaligned or too small */
  LCR_ERR_ALIGNMENT = -4,     /* pointer/stride alignment the kernel requires is not met */
  LCR_ERR_CUDA = -5,          /* cudaGetLastError() != cudaSuccess after launch */
  LCR_ERR_NO_DEVICE = -6      /* no sm_100 device / driver available */
} LcrStatus;

int lcr_version(void);
const char* lcr_error_string(int status);
/* cudaError_t value of the last launch failure seen by this thread (0 if none). */
int lcr_last_cuda_error(void);
/* Number of kernels launched by this library in this process (monotonic; for bench.py's gpu_launches). */
uint64_t lcr_launch_count(void);

/* Tuning switches (LCR_ROI_FWD, LCR_PASTE, LCR_SELECT, LCR_NMS_RESOLVE, ...: the alternative kernels kept for A/B runs and
 * tests, listed in DESIGN.md).  The process environment is read once, on first use; this call overrides a switch at run time
 * (value == NULL: back to the default).  Results never depend on a switch, only which kernel produces them. */
int lcr_set_tuning(const char* key, const char* value);

/* ------------------------------------------------------------------------------------------------
 * a1. Anchors.  Replaces AnchorGenerator.generate_anchors (src/components/anchor_generator.py:13-37).
 * out[(y*w + x)*A + a] = fp32(x*stride, y*stride, x*stride, y*stride) + base[a]   (one fp32 add)
 * base_anchors_host: HOST pointer to A*4 floats (the caller evaluates the float64 sqrt formula of
 * anchor_generator.py:17-27 and rounds to fp32, exactly as torch.tensor(..., dtype=float32) does).
 * ---------------------------------------------------------------------------------------------- */
int lcr_anchors_f32(float* out, int h, int w, int stride, const float* base_anchors_host, int A,
                    void* stream);

/* ------------------------------------------------------------------------------------------------
 * a5/a6. Box clip (in place) and min-size mask.  Replace clip_boxes_to_image / filter_small_boxes
 * (src/utils/box_utils.py:32-37, :39-44).  keep[i] = (x2-x1 >= min_size) & (y2-y1 >= min_size).
 * ---------------------------------------------------------------------------------------------- */
int lcr_clip_boxes_f32(float* boxes, int K, float img_h, float img_w, void* stream);
int lcr_filter_small_boxes_f32(const float* boxes, int K, float min_size, uint8_t* keep, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a7. Delta -> box decode (the inverse of encode_boxes, src/utils/box_utils.py:4-28), pinned to
 * torchvision BoxCoder.decode_single (TV:models/detection/_utils.py:183-224):
 *   w = x2-x1, cx = x1 + 0.5*w; d* = delta/weight; dw,dh clamped to <= xform_clip;
 *   pcx = dx*w + cx; pw = exp(dw)*w; box = (pcx - 0.5*pw, pcy - 0.5*ph, pcx + 0.5*pw, pcy + 0.5*ph)
 * deltas, anchors, out: [K,4].  If img_h > 0 the result is also clipped to [0,img_w]x[0,img_h].
 * ---------------------------------------------------------------------------------------------- */
int lcr_box_decode_f32(const float* deltas, const float* anchors, int K, const float weights_host[4],
                       float xform_clip, float img_h, float img_w, float* out, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a2/a3/a12. RPN proposal selection, batched over images and FPN levels in ONE launch.
 * Replaces the body of generate_inference_proposals / generate_training_proposals up to (not
 * including) NMS (src/utils/proposal_utils.py:16-29, :38-52) and, with topk_on_sigmoid = 0 and
 * deltas given, torchvision's per-level _get_top_n_idx + decode + clip + remove_small + score filter
 * (TV:models/detection/rpn.py:231-297).
 *
 * For every segment s = b*L + l (image b, level l), over the n = A*h*w objectness logits of that
 * level, flat index i = (y*w + x)*A + a  <->  objectness[b][a][y][x]  (the permute(1,2,0).reshape(-1)
 * of proposal_utils.py:16,38):
 *   1. key_i = sigmoid(logit_i) if topk_on_sigmoid else logit_i;    score_i = sigmoid(logit_i)
 *   2. select the k' = min(k, n) largest keys; ties broken by LOWER flat index; NaN ranks highest;
 *      result sorted by (key desc, index asc)                        [torch.topk, proposal_utils.py:19,41]
 *   3. box_i = anchors[i] (gathered, or generated from base_anchors/stride when anchors == NULL),
 *      decoded with deltas[b][4a..4a+3][y][x] when deltas != NULL    [TV rpn.py:366-369]
 *   4. clip to [0,img_w] x [0,img_h]                                 [box_utils.py:32-37]
 *   5. keep, order-preserving, entries with score > thr (>= if !score_strict) and
 *      (x2-x1 >= min_size) & (y2-y1 >= min_size)                     [proposal_utils.py:21-29,43-52]
 * Outputs (capacity k per segment, entries past counts[s] are unspecified):
 *   boxes [B*L, k, 4] f32, scores [B*L, k] f32, index [B*L, k] i64 (flat index), counts [B*L] i32.
 * ---------------------------------------------------------------------------------------------- */
typedef struct LcrRpnLevel {
  const float* objectness;   /* [B, A, h, w] logits */
  const float* deltas;       /* [B, 4A, h, w] or NULL */
  const float* anchors;      /* [h*w*A, 4] or NULL (generate: base_anchors + (x,y)*stride) */
  int h, w, stride;
  int reserved;
  float base_anchors[LCR_MAX_ANCHORS * 4];
} LcrRpnLevel;

typedef struct LcrRpnCfg {
  int num_anchors;           /* A */
  int pre_nms_top_n;         /* k per segment, <= LCR_MAX_TOPK */
  float score_thresh;
  int score_strict;          /* 1: score > thr (reference), 0: score >= thr (torchvision) */
  float min_size;
  int img_h, img_w;
  int topk_on_sigmoid;       /* 1: reference (proposal_utils.py:16-19), 0: torchvision (rpn.py:237) */
  float decode_weights[4];   /* used when deltas != NULL; RPN: (1,1,1,1) */
  float xform_clip;          /* log(1000/16) */
} LcrRpnCfg;

size_t lcr_rpn_select_workspace_bytes(int B, int L, int k);
int lcr_rpn_select_f32(const LcrRpnLevel* levels_host, int L, int B, const LcrRpnCfg* cfg_host,
                       float* boxes, float* scores, int64_t* index, int* counts,
                       void* workspace, size_t workspace_bytes, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a8. Greedy NMS, batched over S independent segments.  Replaces torchvision.ops.nms at
 * src/utils/proposal_utils.py:55 (thr 0.4) and src/custom_maskrcnn.py:192 (thr 0.5)
 * (wrapper TV:ops/boxes.py:20-48, op schema torchvision::nms(Tensor dets, Tensor scores, float thr)).
 *
 * Segment s holds n_s = counts[s] (or `stride` if counts == NULL) boxes at boxes[s*stride ...].
 *   - scores != NULL: candidates are ordered by a STABLE descending sort of scores (ties -> lower
 *     index first, NaN first), after dropping entries with score <= score_thresh when
 *     use_score_thresh != 0 (the `box_scores > 0.4` filter of src/custom_maskrcnn.py:185-188).
 *   - scores == NULL: boxes are taken as already sorted (the rpn_select output order).
 *   - category != NULL: boxes only suppress boxes of the same category (batched_nms,
 *     TV:ops/boxes.py:51-120, per-class branch).
 *   - box j is suppressed by an earlier kept box i iff (double)IoU_fp32(i,j) > iou_threshold, with
 *     IoU = inter / (area_i + area_j - inter), no +1, NaN never suppresses.  This is the comparison of
 *     torchvision's CPU op (fp32 IoU against the double threshold).  Its CUDA op — what the reference
 *     executes on a GPU — rounds the threshold to fp32 first; the two differ exactly when
 *     IoU == (float)thr and (float)thr > thr (0.4f: yes; 0.5f: no), which anchor-grid boxes do produce.
 *     A caller gets the CUDA op's rule by passing (double)(float)thr; the Python drop-ins (`nms`,
 *     RegionPipeline, ops.nms_batched default) do that, `nms_cpu_rule` / cpu_threshold=True do not
 *     (tests/test_gpu_nms.py::test_threshold_rule_matches_torchvision_cuda).
 * keep [S, post_n] i64: kept ORIGINAL in-segment indices in score order, truncated to post_n
 * (the keep_nms[:num_post_nms] of proposal_utils.py:56); keep_counts [S] i32.
 * ---------------------------------------------------------------------------------------------- */
size_t lcr_nms_workspace_bytes(int S, int stride);
int lcr_nms_f32(const float* boxes, const float* scores, const int* category, const int* counts,
                int S, int stride, double iou_threshold, float score_thresh, int use_score_thresh,
                int post_n, int64_t* keep, int* keep_counts,
                void* workspace, size_t workspace_bytes, void* stream);

/* Gather helper used between stages: out_boxes[s][j] = boxes[s][keep[s][j]] (and scores likewise)
 * for j < keep_counts[s]; also emits torchvision-format rois [S*post_n, 5] = (batch_idx, box) with
 * batch_idx = image_of_segment_host? : s, and batch_idx = -1 for padding rows j >= keep_counts[s].
 * out_valid [S*post_n] u8 = (j < keep_counts[s]) feeds lcr_paste_masks_u8.  Any output pointer may
 * be NULL.  (The anchors[keep] / proposals[keep] indexing of
 * proposal_utils.py:56-57 and custom_maskrcnn.py:193-195.) */
int lcr_gather_kept_f32(const float* boxes, const float* scores, const int64_t* keep,
                        const int* keep_counts, int S, int in_stride, int post_n,
                        float* out_boxes, float* out_scores, float* out_rois, uint8_t* out_valid,
                        void* stream);

/* a12, cross-level step of torchvision's RPN (TV:models/detection/rpn.py:258-291, TV:ops/boxes.py:51-120).
 * Input: lcr_rpn_select_f32's per-level outputs boxes [B, L, k, 4], scores [B, L, k], counts [B, L].  Output, per image,
 * the levels' survivors concatenated in level order: cat_boxes / cat_scores / cat_level [B, L*k] (+ zero / -1 padding)
 * and cat_counts [B].  nms_boxes (required with coordinate_trick, optional otherwise): the boxes to run NMS on — with
 * coordinate_trick != 0 each box is offset by level * (max coordinate of the image's boxes + 1) in fp32, as
 * _batched_nms_coordinate_trick does (torchvision's path up to 1000 boxes per image on CPU / 25 000 on CUDA), so a plain
 * lcr_nms_f32 over nms_boxes reproduces batched_nms including its rounding; without it pass cat_level as `category`
 * (_batched_nms_vanilla).  The keep list then indexes cat_boxes (lcr_gather_kept_f32 with post_n = post_nms_top_n). */
int lcr_rpn_concat_levels_f32(const float* boxes, const float* scores, const int* counts, int B, int L, int k,
                              int coordinate_trick, float* cat_boxes, float* nms_boxes, float* cat_scores,
                              int* cat_level, int* cat_counts, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a11. FPN level assignment.  Replaces LevelMapper.__call__ (TV:ops/poolers.py:73-84):
 *   lvl = clamp(floor(lvl0 + log2(sqrt(area)/s0) + eps), k_min, k_max) - k_min       (fp32)
 * boxes [K,4] (or rois [K,5] when box_stride == 5, box at columns 1..4); levels [K] i32.
 * ---------------------------------------------------------------------------------------------- */
int lcr_level_map_f32(const float* boxes, int box_stride, int K, int k_min, int k_max,
                      float canonical_scale, int canonical_level, float eps, int* levels,
                      void* stream);

/* ------------------------------------------------------------------------------------------------
 * a9/a10. RoIAlign forward / backward, single- or multi-level in one launch.
 * Replaces torchvision::roi_align / torchvision::_roi_align_backward as called by
 * RoIAlign(output_size=(7,7), spatial_scale=0.25, sampling_ratio=2) at src/custom_maskrcnn.py:48-50,
 * :120, :177 (module TV:ops/roi_align.py:263-283; kernel math restated in SURVEY.md App. B.1/B.2),
 * and, with L > 1 and roi_level given, MultiScaleRoIAlign (TV:ops/poolers.py:147-227).
 *
 * Each level is a [N, C, H, W] fp32 tensor addressed through ELEMENT strides, so both NCHW and
 * channels_last (NHWC memory) maps are accepted.  Forward dispatch on the fast path: PH == PW == 7, C == 256, K >= 8 x SMs and N >= 4 frames (the
 * reference's pooler on a serving batch) runs the row-program kernel with pipelined rows (roi_fwd_rmp_kernel: results within ~1e-7 of the
 * sample-walk kernel, the taps of a shared window row are pre-added); everything else the sample-walk warp kernel.  A persistent variant
 * (LCR_ROI_FWD=team) uses two words of library-owned device memory per launch (a work counter; 1024 rotating slots, so launches that
 * overlap in time — including replays of different CUDA graphs — do not share one) and marks RoIs for its second pass inside `out`.
 * The warp-item fast paths (pooled tile assembled in shared memory and
 * written by one TMA bulk store; backward: bulk load + vector reductions) need sampling_ratio == 2, PH == PW in {7, 14},
 * sc == 1 (NHWC), C % 4 == 0, even sn / sh / sw, 8-byte aligned maps at least 4 columns wide and a 16-byte aligned
 * out / grad_out.  Every other configuration (NCHW maps as the reference's FPN emits them, any pooled size or sampling ratio)
 * takes roi_fwd_planes_kernel forward — one CTA per (RoI, channel chunk), the RoI's axis taps in shared memory, the arithmetic
 * of torchvision's per-output kernel operation for operation — and the per-output generic kernel backward (same results;
 * LCR_ROI_FWD=generic selects the per-output forward kernel).
 * rois [K,5] = (batch_idx as float, x1, y1, x2, y2) (TV:ops/_utils.py:18-25).  A roi with batch_idx < 0, batch_idx >= N of
 * its level, or a roi_level outside [0, L) is padding: forward writes zeros for it, backward ignores it (no out-of-range
 * access).  roi_level [K] i32 or NULL (all rois on level 0).  out / grad_out: [K, C, PH, PW] contiguous.
 * Backward accumulates into grad levels with the same shapes/strides as the forward features; when zero_grad != 0 the
 * callee zero-fills them first (on `stream`) — that needs DENSE levels (NCHW-contiguous or channels_last strides, so that
 * the level is one run of N*C*H*W floats from `data`): anything else returns LCR_ERR_INVALID_ARG before touching memory.
 * ---------------------------------------------------------------------------------------------- */
/* The `aligned` argument of lcr_roi_align_*_f32 is a flag word:
 *   bit 0  LCR_ROI_ALIGNED     torchvision's `aligned` (half-pixel shift, no minimum RoI size)
 *   bit 1  LCR_ROI_CPU_COORDS  round the sample coordinates as torchvision's CPU kernel does (every operation rounded).
 *          Default (bit clear): as its CUDA kernel does — nvcc contracts `roi * spatial_scale - offset` and
 *          `roi_start + ph * bin_size` into FMAs.  The reference executes the CUDA op on a GPU; torchvision's own two ops
 *          differ by an ulp of the coordinate = up to ~2e-5 of the output range on white-noise features, and this library
 *          agrees with the matching one to ~1.5e-7 (tools/roi_coord_rounding_exp.py).  The CPU-generated golden vectors and
 *          the oracle's default use bit 1. */
#define LCR_ROI_ALIGNED 1
#define LCR_ROI_CPU_COORDS 2

typedef struct LcrFeatLevel {
  float* data;               /* features (forward, read-only) or grad_input (backward, accumulated) */
  int N, H, W;
  int reserved;
  int64_t sn, sc, sh, sw;    /* element strides of the logical [N, C, H, W] view */
  float spatial_scale;
  int reserved2;
} LcrFeatLevel;

int lcr_roi_align_fwd_f32(const LcrFeatLevel* levels_host, int L, int C,
                          const float* rois, const int* roi_level, int K,
                          int PH, int PW, int sampling_ratio, int aligned,
                          float* out, void* stream);
int lcr_roi_align_bwd_f32(const float* grad_out, const LcrFeatLevel* grad_levels_host, int L, int C,
                          const float* rois, const int* roi_level, int K,
                          int PH, int PW, int sampling_ratio, int aligned,
                          int zero_grad, void* stream);

/* Layout helpers (tiled transpose): [N,C,H,W] contiguous <-> NHWC memory ([N,H,W,C] contiguous). */
int lcr_nchw_to_nhwc_f32(const float* in, float* out, int N, int C, int H, int W, void* stream);
int lcr_nhwc_to_nchw_f32(const float* in, float* out, int N, int C, int H, int W, void* stream);

/* ------------------------------------------------------------------------------------------------
 * a13. Batched mask paste / threshold.  Replaces CustomMaskRCNN._generate_masks' per-detection loop
 * (src/custom_maskrcnn.py:276-295) and paste_masks_in_image (src/utils/mask_utils.py:129-171):
 *   (x1,y1,x2,y2) = trunc-to-int(box); x1,y1 = max(0,.); x2 = min(W,.); y2 = min(H,.)
 *   if x2 > x1 and y2 > y1: frame[y1:y2, x1:x2] = bilinear(prob MxM -> (y2-y1, x2-x1),
 *       align_corners=False) > thr ? on_value : 0;   everything else 0.
 * probs [N, M, M] f32 (already sigmoid-ed, class-1 channel), boxes [N,4] f32, out [N, H, W] u8.
 * valid [N] u8 or NULL: frames with valid[i] == 0 are SKIPPED (left untouched) — padded slots of a
 * batched pipeline.  Bilinear follows ATen upsample_bilinear2d.
 * ---------------------------------------------------------------------------------------------- */
int lcr_paste_masks_u8(const float* probs, const float* boxes, const uint8_t* valid, int N, int M,
                       int H, int W, float threshold, uint8_t on_value, uint8_t* out, void* stream);

/* a13, torchvision variant (P2).  Replaces torchvision.models.detection.roi_heads.paste_masks_in_image
 * (TV:models/detection/roi_heads.py:405-501; called by the transfer model's postprocess, src/train_transfer.py:95 ->
 * RoIHeads.forward -> GeneralizedRCNNTransform.postprocess): zero-pad the M x M probability map by `padding`, expand the box
 * by (M + 2*padding)/M around its centre, truncate to integers, resize to (y2-y1+1, x2-x1+1) (bilinear,
 * align_corners=False) and copy the part inside the frame.  out [N, H, W] float32 PROBABILITIES (torchvision returns
 * [N, 1, H, W]), zero outside the box; no threshold.  valid as in lcr_paste_masks_u8. */
int lcr_paste_masks_tv_f32(const float* probs, const float* boxes, const uint8_t* valid, int N, int M,
                           int H, int W, int padding, float* out, void* stream);

/* Detection records for the multi-GPU all-gather: for segment s and slot j,
 * records[s][j] = (x1, y1, x2, y2, score, label) with label = 1.0 for j < counts[s], else all zero.
 * boxes [S, stride, 4], scores [S, stride]; records [S, stride, 6].
 * (labels = ones, src/custom_maskrcnn.py:204.) */
int lcr_pack_records_f32(const float* boxes, const float* scores, const int* counts, int S,
                         int stride, float* records, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Training-side siblings of the region path.
 *
 * Box IoU.  Replaces torchvision.ops.box_iou (TV:ops/boxes.py:308-370) as called at
 * src/components/rpn.py:72, src/custom_maskrcnn.py:221,249, src/utils/mask_utils.py:93:
 *   iou[i][j] = inter / ((area_i + area_j) - inter), fp32, no +1, 0/0 = NaN.
 * lcr_box_iou_f32 writes the [N,G] matrix; lcr_box_iou_max_f32 fuses the `ious.max(dim=1)` that
 * follows every call site (first index of the maximum; NaN propagates and wins) and never
 * materialises the matrix.  boxes [N,4], gt [G,4]; max_iou [N] f32, argmax [N] i64.
 * ---------------------------------------------------------------------------------------------- */
int lcr_box_iou_f32(const float* boxes, int N, const float* gt, int G, float* iou, void* stream);
int lcr_box_iou_max_f32(const float* boxes, int N, const float* gt, int G, float* max_iou,
                        int64_t* argmax, void* stream);

/* Fused matcher.  Replaces the five-op chain that follows box_iou at every training call site:
 *   src/components/rpn.py:72-81      ious.max(dim=1); pos = max >= 0.5; neg = max < 0.3; pos.sum(); neg.sum()
 *   src/custom_maskrcnn.py:221-225   ious.max(dim=1); labels[max_iou >= 0.4] = 1
 *   src/custom_maskrcnn.py:249-251   foreground_mask = max_iou >= 0.4
 * One kernel, no [N,G] matrix.  max_iou [N] f32 / argmax [N] i64 as lcr_box_iou_max_f32 (either may be NULL);
 * pos_mask[i] = max_iou[i] >= pos_thr, neg_mask[i] = max_iou[i] < neg_thr, both [N] u8 (0/1; viewable as bool), compared in
 * fp32 as ATen compares a float tensor with a scalar — a NaN row is in neither mask; counts [2] i32 (device) = the two
 * population counts (zeroed by the call).  The random ± sub-sampling that follows (torch.randperm, rpn.py:84-98) stays with
 * the caller for RNG parity.  G == 0 is LCR_ERR_INVALID_ARG: the reference returns its no-ground-truth loss before matching. */
int lcr_match_boxes_f32(const float* boxes, int N, const float* gt, int G, float pos_thr, float neg_thr,
                        float* max_iou, int64_t* argmax, uint8_t* pos_mask, uint8_t* neg_mask, int* counts,
                        void* stream);

/* Mask targets, batched.  Replaces the per-positive loop over extract_mask_target
 * (src/utils/mask_utils.py:6-46, :110-113): for target k, crop gt_masks[gt_index[k]] (uint8 [H,W]) to
 * the int-truncated, clipped box (x1 in [0,W-1], x2 in [x1+1,W], same for y) and resize bilinearly
 * (align_corners=False) to M x M floats.  gt_index == NULL: target k uses mask k.  An index outside
 * [0,G) yields an all-zero target.  gt_masks [G,H,W] u8, boxes [K,4] f32, out [K,M,M] f32. */
int lcr_mask_targets_f32(const uint8_t* gt_masks, int G, int H, int W, const float* boxes,
                         const int64_t* gt_index, int K, int M, float* out, void* stream);

/* Mask-head tail.  Replaces the F.interpolate(mask_logits, (M,M), bilinear,
 * align_corners=False) that ends CustomMaskHead.forward (src/components/mask_head.py:52-58) and the
 * sigmoid(mask_logits[:, 1]) of _generate_masks (src/custom_maskrcnn.py:273-274): class `cls` only,
 * upsample + sigmoid in one pass.  logits [K, num_classes, m, m] f32 -> probs [K, M, M] f32
 * (the input format of lcr_paste_masks_u8).  m == M skips the resize. */
int lcr_mask_tail_f32(const float* logits, int K, int num_classes, int cls, int m, int M, float* probs,
                      void* stream);

/* Tile stitching.  Replaces calculate_mask_area_in_region (src/visualize.py:106-130)
 * inside filter_detections_by_border_mini_tiles (:174-257): for detection i, total[i] = #{mask > threshold}
 * and, for each of its rectangles r in rects[rect_offsets[i] .. rect_offsets[i+1]) (any number;
 * (x0,y0,x1,y1) half-open, mask coordinates, already clipped to the frame), in_region[r] = the number of
 * those pixels inside r.  Exact integer counts; the float64 fractions and the > mask_threshold decision stay
 * with the caller, as in the reference.  boxes [N,4] (optional): the detection boxes the masks were pasted
 * with — only the box area is scanned.  masks [N,H,W] u8. */
int lcr_mask_region_counts_u8(const uint8_t* masks, int N, int H, int W, const float* boxes,
                              const int* rects, const int* rect_offsets, int threshold, int* total,
                              int* in_region, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* LCR_H_ */

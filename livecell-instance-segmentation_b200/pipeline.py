"""Batched region pipeline — the B200-first form of the per-image loop of
CustomMaskRCNN.forward_inference (src/custom_maskrcnn.py:164-207).

The reference walks the batch in Python and syncs the host 10+ times per image; here every stage
takes the whole batch, results live in fixed-capacity padded device buffers with device-side counts,
and nothing synchronises until the caller asks for dense per-image outputs:

    rpn_select (1 launch, cluster per image) -> nms (3) -> gather (1) -> roi_align (1)
      -> [box head: PyTorch] -> nms with score filter (3) -> gather (1)
      -> [mask head: PyTorch] -> paste (1) -> records (1)

Default hyper-parameters are the reference's (src/utils/proposal_utils.py:33-36,
src/custom_maskrcnn.py:48-50,185,192,292); BASELINE config C3 uses 2000 / 1000 / 500.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional, Tuple

import torch

from . import ops


@dataclass
class RegionConfig:
    sizes: Tuple[float, ...] = (32, 64, 128)
    aspect_ratios: Tuple[float, ...] = (0.5, 1.0, 2.0)
    stride: int = 4                      # hard-coded in the reference (custom_maskrcnn.py:99,159)
    pre_nms_top_n: int = 250
    rpn_score_thresh: float = 0.3
    rpn_nms_thresh: float = 0.4
    post_nms_top_n: int = 50
    min_box_size: float = 10.0
    pooled_size: int = 7
    spatial_scale: float = 0.25
    sampling_ratio: int = 2
    box_score_thresh: float = 0.4
    det_nms_thresh: float = 0.5
    max_detections: Optional[int] = None  # capacity of the detection buffers (default: post_nms_top_n)
    mask_thresh: float = 0.5
    mask_size: int = 28

    @property
    def det_capacity(self) -> int:
        return self.post_nms_top_n if self.max_detections is None else self.max_detections


@dataclass
class Proposals:
    boxes: torch.Tensor      # [B, post_n, 4]
    scores: torch.Tensor     # [B, post_n]
    counts: torch.Tensor     # [B] i32
    rois: torch.Tensor       # [B*post_n, 5], batch index -1 on padding rows


@dataclass
class Detections:
    boxes: torch.Tensor      # [B, D, 4]
    scores: torch.Tensor     # [B, D]
    counts: torch.Tensor     # [B] i32
    index: torch.Tensor      # [B, D] i64: proposal slot of each detection
    valid: torch.Tensor      # [B*D] u8
    masks: Optional[torch.Tensor] = None    # [B*D, H, W] u8 (slots with valid == 0 are untouched)
    records: Optional[torch.Tensor] = None  # [B, D, 6]


class RegionPipeline:
    def __init__(self, cfg: Optional[RegionConfig] = None):
        self.cfg = cfg or RegionConfig()
        self.base = ops.base_anchors(self.cfg.sizes, self.cfg.aspect_ratios)

    # -- stage 1: objectness -> proposals ----------------------------------------------------------
    def proposals(self, objectness: torch.Tensor, image_size) -> Proposals:
        """objectness [B, A, h, w] level-0 logits (cls_scores[0], custom_maskrcnn.py:166)."""
        c = self.cfg
        n = objectness[0].numel()
        k = min(c.pre_nms_top_n, n)
        boxes, scores, _, counts = ops.rpn_select([objectness], k=k, img_size=image_size, score_thresh=c.rpn_score_thresh,
                                                  min_size=c.min_box_size, strides=[c.stride], base=self.base)
        boxes, scores, counts = boxes[:, 0], scores[:, 0], counts[:, 0].contiguous()
        keep, kc = ops.nms_batched(boxes, None, c.rpn_nms_thresh, post_n=c.post_nms_top_n, counts=counts)
        pb, ps, rois = ops.gather_kept(boxes, scores, keep, kc, want_rois=True)
        return Proposals(pb, ps, kc, rois)

    # -- stage 2: RoIAlign ---------------------------------------------------------------------------
    def pool(self, features: torch.Tensor, rois: torch.Tensor) -> torch.Tensor:
        """features: logical [B, C, h, w]; channels_last memory takes the TMA fast path directly."""
        c = self.cfg
        return ops.roi_align_fwd([features], [c.spatial_scale], rois, None, (c.pooled_size, c.pooled_size), c.sampling_ratio, False)

    # -- stage 3: box scores -> detections ---------------------------------------------------------
    def detections(self, props: Proposals, box_scores: torch.Tensor) -> Detections:
        """box_scores [B, post_n] = softmax(cls_logits)[:, 1] per proposal slot
        (custom_maskrcnn.py:182-195: score > 0.4, then NMS 0.5)."""
        c = self.cfg
        D = c.det_capacity
        keep, kc = ops.nms_batched(props.boxes, box_scores, c.det_nms_thresh, post_n=D, counts=props.counts,
                                   score_thresh=c.box_score_thresh)
        db, ds, _, valid = ops.gather_kept(props.boxes, box_scores, keep, kc, want_rois=False, want_valid=True)
        return Detections(db, ds, kc, keep, valid)

    # -- stage 4: mask probabilities -> frames + records -------------------------------------------
    def paste(self, det: Detections, mask_probs: torch.Tensor, image_size, out: Optional[torch.Tensor] = None) -> Detections:
        """mask_probs [B*D, M, M] = sigmoid(mask_logits[:, 1]) per detection slot
        (custom_maskrcnn.py:273-295)."""
        H, W = image_size
        B, D = det.scores.shape
        det.masks = ops.paste_masks(mask_probs.reshape(B * D, mask_probs.shape[-2], mask_probs.shape[-1]), det.boxes.reshape(B * D, 4),
                                    int(H), int(W), self.cfg.mask_thresh, 255, valid=det.valid, out=out)
        det.records = ops.pack_records(det.boxes, det.scores, det.counts)
        return det

    # -- whole path with the PyTorch heads as callables --------------------------------------------
    def infer(self, objectness: torch.Tensor, features: torch.Tensor, image_size,
              box_head: Callable[[torch.Tensor], torch.Tensor], mask_head: Callable[[torch.Tensor], torch.Tensor]):
        """Batched equivalent of forward_inference's loop.  box_head(roi_features [K,C,7,7]) -> class-1
        probabilities [K]; mask_head(roi_features [N,C,7,7]) -> class-1 mask probabilities [N,M,M].
        Returns the reference's output structure: list of dicts(boxes, labels, scores, masks)."""
        c = self.cfg
        B = objectness.shape[0]
        props = self.proposals(objectness, image_size)
        roi_feat = self.pool(features, props.rois)
        box_scores = box_head(roi_feat).reshape(B, c.post_nms_top_n)
        det = self.detections(props, box_scores)
        D = c.det_capacity
        offs = (torch.arange(B, device=det.index.device) * c.post_nms_top_n).unsqueeze(1)
        flat = (det.index.clamp(min=0, max=c.post_nms_top_n - 1) + offs).reshape(-1)
        flat = torch.where(det.valid.bool(), flat, torch.zeros_like(flat))
        probs = mask_head(roi_feat.index_select(0, flat))
        det = self.paste(det, probs, image_size)
        return self.to_predictions(det, image_size)

    @staticmethod
    def to_predictions(det: Detections, image_size):
        """Dense per-image dicts with the reference's dtypes (custom_maskrcnn.py:202-207, 306-314).
        This is the one place that synchronises (reads the counts)."""
        H, W = image_size
        B, D = det.scores.shape
        counts = det.counts.tolist()
        masks = det.masks.reshape(B, D, H, W) if det.masks is not None else None
        preds = []
        for b, n in enumerate(counts):
            preds.append({
                "boxes": det.boxes[b, :n],
                "labels": torch.ones((n,), dtype=torch.long, device=det.boxes.device),
                "scores": det.scores[b, :n],
                "masks": masks[b, :n] if masks is not None else torch.zeros((n, H, W), dtype=torch.uint8, device=det.boxes.device),
            })
        return preds

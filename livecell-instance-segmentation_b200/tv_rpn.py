"""torchvision-RPN semantics on the B200 kernels (SURVEY.md §8 a12 — the transfer model's proposal stage,
src/train_transfer.py:20-37 -> torchvision GeneralizedRCNN -> RegionProposalNetwork).

``filter_proposals`` is the inference path of ``RegionProposalNetwork.forward`` after the head
(TV:models/detection/rpn.py:339-367 with :231-297): anchors (TV:models/detection/anchor_utils.py:58-133 — aspect ratio =
h/w, base anchors rounded), per-level top-k on the LOGITS, BoxCoder.decode, sigmoid, clip, remove_small_boxes,
``score >= score_thresh``, batched_nms over the levels, first post_nms_top_n — as four launches for the whole batch:

    lcr_rpn_select_f32 (all levels, all images) -> lcr_rpn_concat_levels_f32 -> lcr_nms_f32 (3 kernels) -> lcr_gather_kept_f32
"""
from __future__ import annotations

from typing import Optional, Sequence

import torch

from . import ops


def base_anchors(sizes: Sequence[float], aspect_ratios: Sequence[float]) -> torch.Tensor:
    """AnchorGenerator.generate_anchors for one level (TV:models/detection/anchor_utils.py:58-78), fp32 like torchvision:
    h_ratios = sqrt(ratios), w_ratios = 1 / h_ratios, base = round([-w, -h, w, h] / 2)."""
    scales = torch.as_tensor(sizes, dtype=torch.float32)
    ratios = torch.as_tensor(aspect_ratios, dtype=torch.float32)
    h_ratios = torch.sqrt(ratios)
    w_ratios = 1 / h_ratios
    ws = (w_ratios[:, None] * scales[None, :]).view(-1)
    hs = (h_ratios[:, None] * scales[None, :]).view(-1)
    return (torch.stack([-ws, -hs, ws, hs], dim=1) / 2).round()


def uses_coordinate_trick(n_boxes: int, device_type: str = "cuda") -> bool:
    """batched_nms's own switch (TV:ops/boxes.py:83-86): the coordinate trick unless boxes.numel() exceeds 4 000 (CPU) /
    100 000 (CUDA)."""
    return 4 * n_boxes <= (4000 if device_type == "cpu" else 100_000)


def filter_proposals(objectness: Sequence[torch.Tensor], deltas: Sequence[torch.Tensor], image_size, *, sizes, aspect_ratios,
                     pre_nms_top_n: int = 1000, post_nms_top_n: int = 1000, nms_thresh: float = 0.7, score_thresh: float = 0.0,
                     min_size: float = 1e-3, coordinate_trick: Optional[bool] = None, cpu_nms_threshold: bool = False):
    """objectness: per level [B, A, h, w] logits; deltas: per level [B, 4A, h, w]; image_size (H, W) with H, W multiples of
    every level's grid (torchvision pads images to multiples of 32, so stride_h == stride_w = H // h).
    sizes / aspect_ratios: AnchorGenerator's per-level tuples.  coordinate_trick: None = torchvision's rule for a CUDA
    tensor (uses_coordinate_trick); the CPU-generated golden vectors force the CPU rule.
    Returns (boxes [B, post_n, 4], scores [B, post_n], counts [B] i32, level [B, post_n] i32) — padded, device-resident."""
    H, W = int(image_size[0]), int(image_size[1])
    L = len(objectness)
    strides, bases = [], []
    for l, o in enumerate(objectness):
        h, w = o.shape[-2], o.shape[-1]
        sh, sw = H // h, W // w
        if sh != sw:
            raise ValueError(f"level {l}: stride_h {sh} != stride_w {sw} (pad the image to a multiple of 32 as torchvision does)")
        strides.append(sh)
        ar = aspect_ratios[l] if isinstance(aspect_ratios[0], (tuple, list)) else aspect_ratios
        bases.append(base_anchors(sizes[l], ar))
    A = objectness[0].shape[1]
    k = min(int(pre_nms_top_n), max(int(o.shape[1] * o.shape[2] * o.shape[3]) for o in objectness))
    boxes, scores, _, counts = ops.rpn_select(list(objectness), k=k, img_size=(H, W), score_thresh=score_thresh, min_size=min_size,
                                              strides=strides, base=bases, deltas=list(deltas), score_strict=False,
                                              topk_on_sigmoid=False)
    B = boxes.shape[0]
    if coordinate_trick is None:
        coordinate_trick = uses_coordinate_trick(L * k, "cuda")
    cb, nb, cs, cl, cc = ops.rpn_concat_levels(boxes, scores, counts, coordinate_trick)
    post = int(post_nms_top_n)
    keep, kc = ops.nms_batched(nb if coordinate_trick else cb, cs, nms_thresh, post_n=post, counts=cc,
                               category=None if coordinate_trick else cl, cpu_threshold=cpu_nms_threshold)
    ob, osc, _ = ops.gather_kept(cb, cs, keep, kc, want_rois=False)
    lvl = torch.gather(cl, 1, keep.clamp(min=0, max=cl.shape[1] - 1).to(torch.int64)) if cl.numel() else cl
    return ob, osc, kc, lvl

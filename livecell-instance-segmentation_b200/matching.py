"""Anchor matching and ± sampling of the RPN objectness loss on top of the fused matcher.

The reference's RPN loss (src/components/rpn.py:42-123) builds the [anchors x ground truth] IoU matrix, reduces it to a row
max, thresholds it twice, syncs twice for the two mask sums, and then draws the ± sample with torch.randperm.  Here the matrix,
the row max, the masks and the sums are ONE kernel (ops.match_boxes -> lcr_match_boxes_f32, one host sync for both counts); the
random draw stays torch.randperm, called with the same arguments in the same order, so that the same generator state yields the
same sample as the reference — RNG parity is part of the bar, as for sample_proposals.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

from . import ops

POS_IOU, NEG_IOU = 0.5, 0.3          # src/components/rpn.py:76-77
MAX_POS, BATCH = 128, 256            # src/components/rpn.py:80-81


def _draw(mask: torch.Tensor, n: int, device) -> torch.Tensor:
    """n members of `mask` in the order of one torch.randperm over all of them (no draw at all for n == 0, as in
    src/components/rpn.py:84-98, so the generator advances exactly as in the reference)."""
    if n <= 0:
        return torch.tensor([], dtype=torch.long, device=device)
    members = torch.where(mask)[0]
    return members[torch.randperm(len(members), device=device)[:n]]


def sample_rpn_anchors(anchors: torch.Tensor, gt_boxes: torch.Tensor, pos_iou: float = POS_IOU, neg_iou: float = NEG_IOU,
                       max_pos: int = MAX_POS, batch: int = BATCH):
    """src/components/rpn.py:72-105.  Returns (sampled_indices [S] i64, labels [len(anchors)] f32 with 1.0 at the sampled
    positives, pos_sampled, neg_sampled).  Positives come first in sampled_indices, as in the reference."""
    device = anchors.device
    _, _, pos_mask, neg_mask, counts = ops.match_boxes(anchors, gt_boxes, pos_iou, neg_iou)
    n_pos, n_neg = counts.tolist()                       # the one host sync (the reference: two .sum().item())
    num_pos = min(n_pos, max_pos)
    num_neg = min(n_neg, batch - num_pos)
    pos_sampled = _draw(pos_mask, num_pos, device)       # positives first, then negatives: the reference's randperm order
    neg_sampled = _draw(neg_mask, num_neg, device)
    labels = torch.zeros(len(anchors), dtype=torch.float32, device=device)
    labels[pos_sampled] = 1.0
    return torch.cat([pos_sampled, neg_sampled]), labels, pos_sampled, neg_sampled


def rpn_compute_loss(self, cls_scores_list, bbox_deltas_list, anchors, targets, device):
    """Opt-in replacement for RPN.compute_loss (src/components/rpn.py:42-123; install(fused_rpn_matching=True)): same
    branches, same return values, the matching through the fused kernel.  `self` is the reference's RPN module (unused: the loss
    reads no parameters)."""
    cls_scores_flat = cls_scores_list[0].permute(0, 2, 3, 1).reshape(-1)
    all_gt = [t["boxes"] for t in targets if len(t["boxes"]) > 0]
    if len(all_gt) == 0:
        return {"loss_rpn_cls": cls_scores_flat.sum() * 0.0 + 0.1}
    gt_boxes_cat = torch.cat(all_gt)
    if len(gt_boxes_cat) > 0 and len(anchors) > 0:
        sampled, labels, _, _ = sample_rpn_anchors(anchors, gt_boxes_cat)
        if len(sampled) > 0:
            return {"loss_rpn_cls": F.binary_cross_entropy_with_logits(cls_scores_flat[sampled], labels[sampled])}
    return {"loss_rpn_cls": cls_scores_flat.mean() * 0.1}

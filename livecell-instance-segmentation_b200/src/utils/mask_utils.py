"""Drop-in for the reference's ``src/utils/mask_utils.py`` (same names, signatures and return types).

* ``paste_masks_in_image`` (mask_utils.py:129-171) — one batched kernel instead of the per-detection loop.
* ``extract_mask_target`` / ``compute_mask_loss_from_gt`` (mask_utils.py:6-126) — the
  proposal<->GT matching is the fused IoU row-max kernel and the per-positive crop+resize loop (two ``.item()``
  host syncs per proposal in the reference) is one batched kernel; the BCE itself stays on PyTorch.
"""
import torch
import torch.nn.functional as F

from ... import ops


def paste_masks_in_image(masks, boxes, image_size, threshold=0.5):
    """Paste predicted masks [N,M,M] into full frames -> uint8 [N,H,W] in {0,255}
    (mask_utils.py:129-171; same semantics as CustomMaskRCNN._generate_masks' loop,
    custom_maskrcnn.py:276-295).  One kernel, no per-detection host syncs."""
    img_h, img_w = image_size
    return ops.paste_masks(masks, boxes, int(img_h), int(img_w), threshold=float(threshold), on_value=255)


def extract_mask_target(gt_mask, box, mask_size=28):
    """gt_mask [H,W], box [4] -> [mask_size, mask_size] float (mask_utils.py:6-46)."""
    return ops.mask_targets(gt_mask.unsqueeze(0), box.reshape(1, 4), None, int(mask_size))[0]


def compute_mask_loss_from_gt(mask_logits, proposals, targets, device, mask_size=28):
    """Same contract as mask_utils.py:49-126: IoU-match proposals to all GT boxes, keep max IoU > 0.3, BCE between
    the class-1 logits and the bilinear mask targets of the matched GT masks."""
    if len(proposals) == 0 or mask_logits is None:
        return torch.tensor(0.0, device=device)
    gt_boxes = [t["boxes"] for t in targets if len(t["boxes"]) > 0]
    gt_masks = [t["masks"] for t in targets if len(t["boxes"]) > 0]
    if len(gt_boxes) == 0:
        return torch.tensor(0.0, device=device, requires_grad=True)
    gt_boxes, gt_masks = torch.cat(gt_boxes), torch.cat(gt_masks)
    max_ious, matched = ops.box_iou_max(proposals, gt_boxes)                      # mask_utils.py:93-94
    positive = max_ious > 0.3
    if positive.sum() == 0:
        return torch.tensor(0.0, device=device, requires_grad=True)
    idx = matched[positive]
    mask_targets = ops.mask_targets(gt_masks, gt_boxes[idx], idx, int(mask_size))  # mask_utils.py:106-115
    return F.binary_cross_entropy_with_logits(mask_logits[positive][:, 1], mask_targets, reduction="mean")

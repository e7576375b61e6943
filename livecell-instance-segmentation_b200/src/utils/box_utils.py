"""Drop-in for the reference's ``src/utils/box_utils.py``."""
import torch

from ... import ops


def encode_boxes(boxes, anchors):
    """Anchor-relative encoding (reference box_utils.py:4-28).  Loss-side helper, not on the hot
    path: plain torch ops, any device.  Its inverse on the hot path is ops.box_decode."""
    anchors_w = (anchors[:, 2] - anchors[:, 0]).clamp(min=1.0)
    anchors_h = (anchors[:, 3] - anchors[:, 1]).clamp(min=1.0)
    boxes_w = (boxes[:, 2] - boxes[:, 0]).clamp(min=1.0)
    boxes_h = (boxes[:, 3] - boxes[:, 1]).clamp(min=1.0)
    acx, acy = (anchors[:, 0] + anchors[:, 2]) / 2.0, (anchors[:, 1] + anchors[:, 3]) / 2.0
    bcx, bcy = (boxes[:, 0] + boxes[:, 2]) / 2.0, (boxes[:, 1] + boxes[:, 3]) / 2.0
    return torch.stack([(bcx - acx) / anchors_w, (bcy - acy) / anchors_h,
                        torch.log(boxes_w / anchors_w), torch.log(boxes_h / anchors_h)], dim=1)


def decode_boxes(deltas, anchors, weights=(1.0, 1.0, 1.0, 1.0), image_size=None):
    """Delta -> box decode (the kernel the north star asks for; the reference itself never decodes)."""
    return ops.box_decode(deltas, anchors, weights, img_size=image_size)


def clip_boxes_to_image(boxes, image_size):
    """Clip boxes to image boundaries IN PLACE and return the same tensor (box_utils.py:32-37)."""
    h, w = image_size
    return ops.clip_boxes_(boxes, h, w)


def filter_small_boxes(boxes, min_size=1):
    """bool[K] mask of boxes with w >= min_size and h >= min_size (box_utils.py:39-44)."""
    return ops.filter_small_boxes(boxes, min_size)

"""Drop-in for the paste half of the reference's ``src/utils/mask_utils.py``.

``extract_mask_target`` / ``compute_mask_loss_from_gt`` (mask_utils.py:6-126) are loss-side and listed
as "next" in SURVEY.md §8(f); they are not part of this package yet — install() leaves the
reference's own implementations in place."""
from ... import ops


def paste_masks_in_image(masks, boxes, image_size, threshold=0.5):
    """Paste predicted masks [N,M,M] into full frames -> uint8 [N,H,W] in {0,255}
    (mask_utils.py:129-171; same semantics as CustomMaskRCNN._generate_masks' loop,
    custom_maskrcnn.py:276-295).  One kernel, no per-detection host syncs."""
    img_h, img_w = image_size
    return ops.paste_masks(masks, boxes, int(img_h), int(img_w), threshold=float(threshold), on_value=255)

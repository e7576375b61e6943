"""install() — swap the region-pipeline entry points of an importable reference checkout for the
B200-native ones, so that ``train_custom.py`` / ``app_gradio.py`` run unchanged.

The reference binds its hot-path callables by name at import time
(``src/custom_maskrcnn.py:5,12-19``); install() rebinds those names in the already-imported (or
freshly imported) reference modules.  Nothing in the reference tree is modified on disk.
"""
from __future__ import annotations

import importlib
import sys
import types

from . import roi_align as _ra
from .src.components import anchor_generator as _ag
from .src.utils import box_utils as _bu
from .src.utils import mask_utils as _mu
from .src.utils import proposal_utils as _pu

# reference module name -> {attribute: replacement}
PATCHES = {
    "src.components.anchor_generator": {"AnchorGenerator": _ag.AnchorGenerator},
    "src.utils.box_utils": {"clip_boxes_to_image": _bu.clip_boxes_to_image, "filter_small_boxes": _bu.filter_small_boxes},
    "src.utils.proposal_utils": {
        "generate_training_proposals": _pu.generate_training_proposals,
        "generate_inference_proposals": _pu.generate_inference_proposals,
        "sample_proposals": _pu.sample_proposals,
        "nms": _ra.nms,
        "clip_boxes_to_image": _bu.clip_boxes_to_image,
        "filter_small_boxes": _bu.filter_small_boxes,
    },
    "src.utils.mask_utils": {"paste_masks_in_image": _mu.paste_masks_in_image, "extract_mask_target": _mu.extract_mask_target,
                             "compute_mask_loss_from_gt": _mu.compute_mask_loss_from_gt, "box_iou": _ra.box_iou},
    "src.components.rpn": {"box_iou": _ra.box_iou},
    "src.custom_maskrcnn": {
        "AnchorGenerator": _ag.AnchorGenerator,
        "RoIAlign": _ra.RoIAlign,
        "nms": _ra.nms,
        "box_iou": _ra.box_iou,
        "compute_mask_loss_from_gt": _mu.compute_mask_loss_from_gt,
        "generate_training_proposals": _pu.generate_training_proposals,
        "generate_inference_proposals": _pu.generate_inference_proposals,
        "sample_proposals": _pu.sample_proposals,
    },
}
# scripts do sys.path.append('src') and `from custom_maskrcnn import ...` (src/train_custom.py:15-16)
ALIASES = {"custom_maskrcnn": "src.custom_maskrcnn"}


def _paste_method(self, roi_features, boxes, image_size, device):
    """Replacement for CustomMaskRCNN._generate_masks (src/custom_maskrcnn.py:265-295): the mask head
    stays on PyTorch, the per-detection paste loop becomes one kernel."""
    import torch
    img_h, img_w = image_size
    if len(boxes) == 0:
        return torch.zeros((0, img_h, img_w), dtype=torch.uint8, device=device)
    mask_probs = mask_head_probs(self.mask_head, roi_features)
    return _mu.paste_masks_in_image(mask_probs, boxes, (img_h, img_w), threshold=0.5)


_HEAD_LAYERS = ("conv1", "conv2", "conv3", "conv4", "deconv", "deconv_relu", "mask_fcn_logits")


def mask_head_probs(mask_head, roi_features):
    """sigmoid(mask_head(x)[:, 1]) with the head's tail fused: the reference head's own layers
    (src/components/mask_head.py:41-50, weights untouched) produce the 14x14 logits; the final bilinear 14->28 of both
    classes + class select + sigmoid (mask_head.py:52-58, custom_maskrcnn.py:273-274) become one kernel over class 1.
    A head without those attributes is simply called, and only the sigmoid/select is fused."""
    from . import ops
    if all(hasattr(mask_head, n) for n in _HEAD_LAYERS) and hasattr(mask_head, "mask_size"):
        x = roi_features
        for n in _HEAD_LAYERS:
            x = getattr(mask_head, n)(x)
        return ops.mask_tail(x, int(mask_head.mask_size), 1)
    logits = mask_head(roi_features)
    return ops.mask_tail(logits, int(logits.shape[-1]), 1)


_ORIGINALS: list = []   # (owner object, attribute, original value) of everything install() replaced, in order


def _swap(owner, attr, obj):
    _ORIGINALS.append((owner, attr, getattr(owner, attr)))
    setattr(owner, attr, obj)


def uninstall() -> int:
    """Undo every install() since the last uninstall(): the reference modules get their own callables back
    (models constructed while installed keep their B200 RoIAlign module).  Returns the number of names restored."""
    n = len(_ORIGINALS)
    while _ORIGINALS:
        owner, attr, orig = _ORIGINALS.pop()
        setattr(owner, attr, orig)
    return n


def _forward_inference_batched(self, images):
    """Opt-in replacement for CustomMaskRCNN.forward_inference (src/custom_maskrcnn.py:144-209): the same computation with the
    per-image Python loop (about ten host syncs per image) replaced by the batched region pipeline — proposals, RoIAlign, box
    head, score filter + NMS, mask head and paste each run ONCE for the whole batch; backbone, FPN, RPN head, box head and mask
    head are the model's own modules.  Returns the reference's structure: one dict(boxes, labels, scores, masks) per image,
    detections in NMS (score) order.  Hyper-parameters are the reference's defaults (proposal_utils.py:33-36,
    custom_maskrcnn.py:185,192)."""
    import torch.nn.functional as F
    from .pipeline import RegionConfig, RegionPipeline
    features, images_tensor = self.extract_features(images)
    cls_scores, _ = self.rpn(features)
    pipe = getattr(self, "_lcr_region_pipeline", None)
    if pipe is None:
        pipe = RegionPipeline(RegionConfig())
        object.__setattr__(self, "_lcr_region_pipeline", pipe)      # not a parameter / buffer / sub-module: checkpoints unchanged
    return pipe.infer(cls_scores[0], features[0], tuple(int(v) for v in images_tensor.shape[-2:]),
                      box_head=lambda rf: F.softmax(self.box_head(rf)[0], dim=-1)[:, 1],
                      mask_head=lambda rf: mask_head_probs(self.mask_head, rf))


def install(import_missing: bool = True, batched_inference: bool = False, fused_rpn_matching: bool = False) -> dict:
    """Patch every reference module that is (or can be) imported.  Returns {module: [patched names]}.
    batched_inference=True additionally swaps CustomMaskRCNN.forward_inference for the batched region pipeline
    (_forward_inference_batched): same results, no per-image loop.  fused_rpn_matching=True swaps RPN.compute_loss
    (src/components/rpn.py:42-123) for matching.rpn_compute_loss: IoU matrix, row max, threshold masks and their sums in one
    kernel, the same torch.randperm draws (same sample for the same generator state)."""
    done = {}
    targets = []                       # resolve (import) every module first, so that each one still binds the reference's own
    for mod_name, repl in PATCHES.items():   # callables when it is patched and uninstall() can give them back
        for name in [mod_name] + [a for a, target in ALIASES.items() if target == mod_name]:
            mod = sys.modules.get(name)
            if mod is None and import_missing:
                try:
                    mod = importlib.import_module(name)
                except Exception:
                    mod = None
            if isinstance(mod, types.ModuleType):
                targets.append((name, mod, repl))
    for name, mod, repl in targets:
        for attr, obj in repl.items():
            if hasattr(mod, attr):
                if getattr(mod, attr) is not obj:
                    _swap(mod, attr, obj)
                done.setdefault(name, []).append(attr)
        if fused_rpn_matching and name == "src.components.rpn" and hasattr(mod, "RPN"):
            from .matching import rpn_compute_loss
            if mod.RPN.compute_loss is not rpn_compute_loss:
                _swap(mod.RPN, "compute_loss", rpn_compute_loss)
            done.setdefault(name, []).append("RPN.compute_loss")
        if name.endswith("custom_maskrcnn") and hasattr(mod, "CustomMaskRCNN"):
            if mod.CustomMaskRCNN._generate_masks is not _paste_method:
                _swap(mod.CustomMaskRCNN, "_generate_masks", _paste_method)
            done.setdefault(name, []).append("CustomMaskRCNN._generate_masks")
            if batched_inference:
                if mod.CustomMaskRCNN.forward_inference is not _forward_inference_batched:
                    _swap(mod.CustomMaskRCNN, "forward_inference", _forward_inference_batched)
                done[name].append("CustomMaskRCNN.forward_inference")
    return done

"""Drop-ins for the three torchvision operators the reference imports
(``from torchvision.ops import RoIAlign, nms, box_iou`` — src/custom_maskrcnn.py:5):

* ``RoIAlign`` — parameter-free nn.Module (checkpoints interchange; count_parameters keeps reporting
  ``roi_align: 0``), differentiable w.r.t. the feature map through ``_RoIAlignFn``.
* ``nms`` / ``batched_nms`` — same signatures and return type (int64 kept indices, score order).
* ``MultiScaleRoIAlign`` — the multi-level pooler of the transfer model (TV:ops/poolers.py:230-321),
  level assignment + all levels in ONE launch.
* ``box_iou`` / ``box_iou_max`` — the IoU matrix in one kernel, and the fused ``box_iou(a, b).max(dim=1)``
  every loss-side call site computes next.
"""
from __future__ import annotations

from typing import Sequence, Union

import torch
from torch import nn

from . import _ext, _lib, ops


def convert_boxes_to_roi_format(boxes: Sequence[torch.Tensor]) -> torch.Tensor:
    """list[Tensor[K_i,4]] -> Tensor[K,5] with the image index in column 0 (TV:ops/_utils.py:18-25)."""
    if len(boxes) == 1:   # the reference's call form (feature_map[b:b+1], [proposals]): one pad instead of full + 3 cats
        return torch.nn.functional.pad(boxes[0], (1, 0), value=0.0)
    ids = [torch.full((b.shape[0], 1), float(i), dtype=b.dtype, device=b.device) for i, b in enumerate(boxes)]
    return torch.cat([torch.cat(ids, dim=0), torch.cat(list(boxes), dim=0)], dim=1)


def _as_rois(rois) -> torch.Tensor:
    if isinstance(rois, (list, tuple)):
        return convert_boxes_to_roi_format(rois)
    return rois


def _prefer_nhwc(x: torch.Tensor, K: int, PH: int, PW: int) -> bool:
    """NCHW-contiguous input: transpose the map once (tiled 16-byte transpose) and pool with the NHWC warp-item kernels, or
    pool the planes directly (roi_fwd_planes_kernel)?  Measured on B200 (profiles/r02c_roi_nchw_exp.jsonl, kernel time inside
    a CUDA graph, 256 channels, 7x7): the planes kernel costs ~0.2 us per RoI (LSU-wavefront bound: a warp's 16 loads per
    output each touch ~6 lines), the transpose ~15 us per 130x176 map plus the NHWC kernel's ~6 us floor: the direct kernel
    wins below ~100 RoIs on a 130x176 map and ~70 on a 64x64 one.  In elements: pooled output >= 8e5 + 8 % of the map."""
    N, C, H, W = x.shape
    return K * PH * PW * C >= 800_000 + 0.08 * N * C * H * W


class _RoIAlignFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, feats_and_meta, *feats):
        scales, rois, roi_level, output_size, sampling_ratio, aligned = feats_and_meta
        PH, PW = output_size
        K = rois.shape[0]
        used = []
        for f in feats:
            N, C, H, W = f.shape
            nhwc_dense = f.dtype == torch.float32 and f.stride() == (H * W * C, 1, W * C, C)
            if not nhwc_dense and sampling_ratio == 2 and PH == PW and PH in (7, 14) and C % 4 == 0 and _prefer_nhwc(f, K, PH, PW):
                f = ops.to_nhwc(f)
            used.append(f)
        out = ops.roi_align_fwd(used, scales, rois, roi_level, (PH, PW), sampling_ratio, aligned)
        ctx.meta = (scales, output_size, sampling_ratio, aligned, [tuple(f.shape) for f in feats])
        if roi_level is None:
            ctx.save_for_backward(rois)
        else:
            ctx.save_for_backward(rois, roi_level)
        ctx.has_level = roi_level is not None
        return out

    @staticmethod
    def backward(ctx, grad_out):
        scales, output_size, sampling_ratio, aligned, shapes = ctx.meta
        rois = ctx.saved_tensors[0]
        roi_level = ctx.saved_tensors[1] if ctx.has_level else None
        grads = []
        for (N, C, H, W) in shapes:
            # channels_last grad buffers: the backward fast path scatters whole channel vectors
            grads.append(torch.empty_strided((N, C, H, W), (H * W * C, 1, W * C, C), dtype=torch.float32, device=grad_out.device))
        ops.roi_align_bwd(grad_out, grads, scales, rois, roi_level, sampling_ratio, aligned, zero_grad=True)
        return (None, *grads)


def roi_align(input: torch.Tensor, boxes, output_size, spatial_scale: float = 1.0, sampling_ratio: int = -1,
              aligned: bool = False) -> torch.Tensor:
    """torchvision.ops.roi_align signature (TV:ops/roi_align.py:204-260)."""
    if isinstance(output_size, int):
        output_size = (output_size, output_size)
    sr = max(int(sampling_ratio), 0)  # torchvision: <= 0 means adaptive
    if not input.is_cuda:
        raise _lib.LcrError("liblcr ops need CUDA tensors: the region pipeline has no CPU fallback")
    ext = _ext.load()
    if ext is not None:  # C++ call path + C++ autograd node (csrc/torch_ext.cpp): same kernels, no Python plumbing
        flags = ops.roi_align_flags(bool(aligned))       # bit 0 aligned, bit 1 CPU-op coordinate rounding (ops.set_roi_coord_rule)
        if isinstance(boxes, (list, tuple)):
            return ext.roi_align_list(input, list(boxes), float(spatial_scale), output_size[0], output_size[1], sr, flags)
        return ext.roi_align(input, boxes, float(spatial_scale), output_size[0], output_size[1], sr, flags)
    rois = _as_rois(boxes)
    return _RoIAlignFn.apply(((float(spatial_scale),), rois, None, tuple(output_size), sr, bool(aligned)), input)


class RoIAlign(nn.Module):
    """torchvision.ops.RoIAlign drop-in (TV:ops/roi_align.py:263-283) as constructed at
    src/custom_maskrcnn.py:48-50: RoIAlign(output_size=(7,7), spatial_scale=0.25, sampling_ratio=2).
    No parameters, no buffers."""

    def __init__(self, output_size, spatial_scale: float, sampling_ratio: int, aligned: bool = False):
        super().__init__()
        self.output_size = output_size
        self.spatial_scale = spatial_scale
        self.sampling_ratio = sampling_ratio
        self.aligned = aligned

    def forward(self, input: torch.Tensor, rois: Union[torch.Tensor, Sequence[torch.Tensor]]) -> torch.Tensor:
        return roi_align(input, rois, self.output_size, self.spatial_scale, self.sampling_ratio, self.aligned)

    def __repr__(self) -> str:
        return (f"{self.__class__.__name__}(output_size={self.output_size}, spatial_scale={self.spatial_scale}, "
                f"sampling_ratio={self.sampling_ratio}, aligned={self.aligned})")


class MultiScaleRoIAlign(nn.Module):
    """Multi-level pooler with torchvision semantics (TV:ops/poolers.py:230-321): scales inferred as
    2^-k from the feature/image sizes, LevelMapper(k_min, k_max, 224, 4), one launch for all levels."""

    def __init__(self, featmap_names, output_size, sampling_ratio, *, canonical_scale: int = 224, canonical_level: int = 4):
        super().__init__()
        self.featmap_names = list(featmap_names)
        self.output_size = (output_size, output_size) if isinstance(output_size, int) else tuple(output_size)
        self.sampling_ratio = sampling_ratio
        self.canonical_scale = canonical_scale
        self.canonical_level = canonical_level

    @staticmethod
    def infer_scales(feats, image_shapes):
        import math
        max_h = max(s[0] for s in image_shapes)
        scales = []
        for f in feats:
            approx = float(f.shape[-2]) / float(max_h)
            scales.append(2.0 ** float(round(math.log2(approx))))
        return scales

    def forward(self, x, boxes, image_shapes):
        feats = [x[k] for k in self.featmap_names] if isinstance(x, dict) else list(x)
        rois = _as_rois(boxes)
        scales = self.infer_scales(feats, image_shapes)
        import math
        k_min, k_max = int(-math.log2(scales[0])), int(-math.log2(scales[-1]))
        levels = ops.level_map(rois, k_min, k_max, float(self.canonical_scale), self.canonical_level) if len(feats) > 1 else None
        sr = max(int(self.sampling_ratio), 0)
        return _RoIAlignFn.apply((tuple(scales), rois, levels, self.output_size, sr, False), *feats)


def nms(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """torchvision.ops.nms drop-in (TV:ops/boxes.py:20-48): int64 indices of the kept boxes, sorted by
    decreasing score.  One host sync (the dense return type needs the count).

    The threshold is rounded to fp32 first, as torchvision's CUDA op does (the op the reference runs on a GPU);
    nms_cpu_rule() keeps the CPU op's double comparison (see ops.nms_batched)."""
    return _nms(boxes, scores, iou_threshold, False)


def nms_cpu_rule(boxes: torch.Tensor, scores: torch.Tensor, iou_threshold: float) -> torch.Tensor:
    """nms() with torchvision's CPU-op threshold rule (fp32 IoU compared with the double threshold): differs from nms()
    only for a pair whose IoU equals float32(iou_threshold) exactly.  The CPU-generated golden vectors follow this rule."""
    return _nms(boxes, scores, iou_threshold, True)


def _nms(boxes, scores, iou_threshold, cpu_threshold):
    if not boxes.is_cuda:
        raise _lib.LcrError("liblcr ops need CUDA tensors: the region pipeline has no CPU fallback")
    thr = float(iou_threshold) if cpu_threshold else ops._round_f32(iou_threshold)
    ext = _ext.load()
    if ext is not None:
        return ext.nms(boxes, scores, thr)
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    keep, kc = ops.nms_batched(boxes.reshape(1, n, 4), scores.reshape(1, n), thr, post_n=n, cpu_threshold=True)
    return keep[0, : int(kc.item())].clone()


def batched_nms(boxes: torch.Tensor, scores: torch.Tensor, idxs: torch.Tensor, iou_threshold: float,
                cpu_threshold: bool = False) -> torch.Tensor:
    """torchvision.ops.batched_nms drop-in (TV:ops/boxes.py:51-120), per-category semantics (boxes of
    different categories never interact), evaluated exactly — no coordinate-offset rounding."""
    n = boxes.shape[0]
    if n == 0:
        return torch.empty((0,), dtype=torch.int64, device=boxes.device)
    keep, kc = ops.nms_batched(boxes.reshape(1, n, 4), scores.reshape(1, n), float(iou_threshold), post_n=n,
                               category=idxs.reshape(1, n), cpu_threshold=cpu_threshold)
    return keep[0, : int(kc.item())].clone()


def box_iou(boxes1: torch.Tensor, boxes2: torch.Tensor) -> torch.Tensor:
    """torchvision.ops.box_iou drop-in (TV:ops/boxes.py:308-370): [N,4] x [M,4] -> [N,M]."""
    return ops.box_iou(boxes1, boxes2)


def match_boxes(boxes1: torch.Tensor, boxes2: torch.Tensor, pos_thr: float, neg_thr=None):
    """``box_iou(boxes1, boxes2).max(dim=1)`` + ``>= pos_thr`` / ``< neg_thr`` masks + their sums, one kernel
    (src/components/rpn.py:72-81, src/custom_maskrcnn.py:221-225): (max_iou, argmax, pos, neg, counts[2])."""
    return ops.match_boxes(boxes1, boxes2, pos_thr, neg_thr)


def box_iou_max(boxes1: torch.Tensor, boxes2: torch.Tensor):
    """(values, indices) of ``box_iou(boxes1, boxes2).max(dim=1)`` without materialising the matrix."""
    return ops.box_iou_max(boxes1, boxes2)

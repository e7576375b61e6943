"""Tensor-level wrappers over the C-ABI (include/lcr.h): torch is used for device memory and the
current stream only.  Every function launches hand-written sm_100a kernels from csrc/liblcr.so on
``torch.cuda.current_stream()``; none of them synchronises.  CPU tensors are rejected loudly.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional, Sequence

import torch

from . import _lib
from ._lib import LcrFeatLevel, LcrRpnCfg, LcrRpnLevel, check

XFORM_CLIP = math.log(1000.0 / 16)

# RoIAlign sample-coordinate rounding: "cuda" (default) = as torchvision's CUDA op, the op the reference executes on a GPU
# (nvcc contracts `roi_start + ph * bin_size` into an FMA); "cpu" = as torchvision's CPU op (every operation rounded), the rule
# of the CPU-generated golden vectors and of the oracle.  torchvision's own two ops differ by up to ~2e-5 of the output range
# on white-noise features; with the matching rule this library agrees with either to ~1.5e-7 (tools/roi_coord_rounding_exp.py).
_roi_cpu_coords = False


def set_roi_coord_rule(rule: str) -> str:
    """'cuda' (default) or 'cpu'; returns the previous rule."""
    global _roi_cpu_coords
    if rule not in ("cuda", "cpu"):
        raise ValueError("rule must be 'cuda' or 'cpu'")
    prev = "cpu" if _roi_cpu_coords else "cuda"
    _roi_cpu_coords = rule == "cpu"
    return prev


def roi_align_flags(aligned: bool, cpu_coords: Optional[bool] = None) -> int:
    """The `aligned` flag word of lcr_roi_align_*_f32: bit 0 aligned, bit 1 CPU-op coordinate rounding."""
    cc = _roi_cpu_coords if cpu_coords is None else bool(cpu_coords)
    return (1 if aligned else 0) | (2 if cc else 0)


_workspaces: dict = {}
_retired: list = []      # outgrown workspaces: kept alive, a CUDA graph captured earlier may still point into them


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


class _NullCtx:
    __slots__ = ()

    def __enter__(self):
        return None

    def __exit__(self, *exc):
        return False


_NULL = _NullCtx()


def _dev(device):
    """Device guard that costs nothing when `device` is already current (the common single-GPU-per-process case);
    torch.cuda.device() alone is ~10 us per call, which matters for the small-K training shapes."""
    device = torch.device(device)
    idx = device.index
    if idx is None or idx == torch.cuda.current_device():
        return _NULL
    return torch.cuda.device(device)


def _need_cuda(*tensors):
    for t in tensors:
        if t is not None and not t.is_cuda:
            raise _lib.LcrError("liblcr ops need CUDA tensors: the region pipeline has no CPU fallback")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.float()
    return t.contiguous()


def _workspace(nbytes: int, device) -> torch.Tensor:
    """Per (device, stream) scratch buffer, grown on demand; owned by the caller side of the ABI."""
    key = (torch.device(device).index, _stream())
    ws = _workspaces.get(key)
    if ws is None or ws.numel() < nbytes:
        if ws is not None:
            _retired.append(ws)   # never handed back to the allocator: graph replays would otherwise write into foreign memory
        ws = torch.empty(max(nbytes, 2 * ws.numel() if ws is not None else 0, 1 << 20), dtype=torch.uint8, device=device)
        _workspaces[key] = ws
    return ws


def _round_f32(x: float) -> float:
    """x rounded to the nearest fp32, as a Python float."""
    import struct
    return struct.unpack("f", struct.pack("f", float(x)))[0]


def base_anchors(sizes=(32, 64, 128), aspect_ratios=(0.5, 1.0, 2.0)):
    """Base anchors in float64 then fp32 — src/components/anchor_generator.py:15-27."""
    rows = []
    for size in sizes:
        for ratio in aspect_ratios:
            area = size * size
            h = math.sqrt(area / ratio)
            w = h * ratio
            rows.append([-w / 2, -h / 2, w / 2, h / 2])
    return torch.tensor(rows, dtype=torch.float64).to(torch.float32)


def _base_array(base) -> "C.Array":
    flat = [float(v) for v in torch.as_tensor(base, dtype=torch.float32).reshape(-1).tolist()]
    if len(flat) > _lib.LCR_MAX_ANCHORS * 4 or len(flat) % 4:
        raise _lib.LcrError("base anchors: at most %d anchors of 4 floats" % _lib.LCR_MAX_ANCHORS)
    return (C.c_float * len(flat))(*flat)


# ------------------------------------------------------------------------------------------------
def anchors(h: int, w: int, stride: int, base, device) -> torch.Tensor:
    """a1 — AnchorGenerator.generate_anchors (src/components/anchor_generator.py:13-37)."""
    device = torch.device(device)
    if device.type != "cuda":
        raise _lib.LcrError("liblcr ops need a CUDA device: the region pipeline has no CPU fallback")
    arr = _base_array(base)
    A = len(arr) // 4
    out = torch.empty((h * w * A, 4), dtype=torch.float32, device=device)
    with _dev(device):
        check(_lib.load().lcr_anchors_f32(out.data_ptr(), h, w, stride, arr, A, _stream()), "anchors")
    return out


def clip_boxes_(boxes: torch.Tensor, img_h: float, img_w: float) -> torch.Tensor:
    """a5 — clip_boxes_to_image, in place (src/utils/box_utils.py:32-37)."""
    _need_cuda(boxes)
    if boxes.numel() == 0:
        return boxes
    if boxes.dtype != torch.float32 or not boxes.is_contiguous():
        tmp = _f32c(boxes)
        clip_boxes_(tmp, img_h, img_w)
        boxes.copy_(tmp)
        return boxes
    with _dev(boxes.device):
        check(_lib.load().lcr_clip_boxes_f32(boxes.data_ptr(), boxes.shape[0], float(img_h), float(img_w), _stream()), "clip")
    return boxes


def filter_small_boxes(boxes: torch.Tensor, min_size: float) -> torch.Tensor:
    """a6 — filter_small_boxes (src/utils/box_utils.py:39-44) -> bool[K]."""
    _need_cuda(boxes)
    b = _f32c(boxes)
    keep = torch.empty((b.shape[0],), dtype=torch.uint8, device=b.device)
    if b.shape[0]:
        with _dev(b.device):
            check(_lib.load().lcr_filter_small_boxes_f32(b.data_ptr(), b.shape[0], float(min_size), keep.data_ptr(), _stream()),
                  "filter_small_boxes")
    return keep.view(torch.bool)


def box_decode(deltas: torch.Tensor, anc: torch.Tensor, weights=(1.0, 1.0, 1.0, 1.0), xform_clip: float = XFORM_CLIP,
               img_size=None) -> torch.Tensor:
    """a7 — BoxCoder.decode_single (TV:models/detection/_utils.py:183-224), optional clip."""
    _need_cuda(deltas, anc)
    d, a = _f32c(deltas).reshape(-1, 4), _f32c(anc).reshape(-1, 4)
    out = torch.empty_like(a)
    w = (C.c_float * 4)(*[float(v) for v in weights])
    ih, iw = (0.0, 0.0) if img_size is None else (float(img_size[0]), float(img_size[1]))
    if a.shape[0]:
        with _dev(a.device):
            check(_lib.load().lcr_box_decode_f32(d.data_ptr(), a.data_ptr(), a.shape[0], w, float(xform_clip), ih, iw,
                                                 out.data_ptr(), _stream()), "box_decode")
    return out


def rpn_select(objectness: Sequence[torch.Tensor], *, k: int, img_size, score_thresh: float, min_size: float,
               strides: Optional[Sequence[int]] = None, base=None, anchors_per_level: Optional[Sequence[torch.Tensor]] = None,
               deltas: Optional[Sequence[torch.Tensor]] = None, score_strict: bool = True, topk_on_sigmoid: bool = True,
               decode_weights=(1.0, 1.0, 1.0, 1.0), xform_clip: float = XFORM_CLIP):
    """a2/a3/a12 — batched proposal selection (see lcr_rpn_select_f32 in include/lcr.h).

    objectness: list over levels of [B, A, h, w] logits.  Boxes come from `anchors_per_level`
    ([h*w*A, 4] tensors) or are generated from `base` (one [A,4] table, or one per level) + `strides`.
    Returns boxes [B, L, k, 4], scores [B, L, k], index [B, L, k] (i64), counts [B, L] (i32)."""
    L = len(objectness)
    objs = [_f32c(o) for o in objectness]
    _need_cuda(*objs)
    B, A = objs[0].shape[0], objs[0].shape[1]
    dev = objs[0].device
    dls = [None] * L if deltas is None else [_f32c(d) for d in deltas]
    ancs = [None] * L if anchors_per_level is None else [_f32c(a) for a in anchors_per_level]
    _need_cuda(*[t for t in dls + ancs if t is not None])
    levels = (LcrRpnLevel * L)()
    for l in range(L):
        lv = levels[l]
        lv.objectness, lv.deltas, lv.anchors = _ptr(objs[l]), _ptr(dls[l]), _ptr(ancs[l])
        lv.h, lv.w = objs[l].shape[2], objs[l].shape[3]
        lv.stride = int(strides[l]) if strides is not None else 0
        if ancs[l] is None:
            if base is None or strides is None:
                raise _lib.LcrError("rpn_select: give anchors_per_level, or base + strides")
            bl = base[l] if isinstance(base, (list, tuple)) else base
            arr = _base_array(bl)
            if len(arr) != A * 4:
                raise _lib.LcrError("rpn_select: base anchors do not match the objectness channel count")
            for i, v in enumerate(arr):
                lv.base_anchors[i] = v
        elif ancs[l].shape[0] != A * lv.h * lv.w:
            raise _lib.LcrError("rpn_select: anchors tensor does not match the objectness map")
    cfg = LcrRpnCfg()
    cfg.num_anchors, cfg.pre_nms_top_n = A, int(k)
    cfg.score_thresh, cfg.score_strict = float(score_thresh), 1 if score_strict else 0
    cfg.min_size = float(min_size)
    cfg.img_h, cfg.img_w = int(img_size[0]), int(img_size[1])
    cfg.topk_on_sigmoid = 1 if topk_on_sigmoid else 0
    for i in range(4):
        cfg.decode_weights[i] = float(decode_weights[i])
    cfg.xform_clip = float(xform_clip)
    boxes = torch.empty((B, L, k, 4), dtype=torch.float32, device=dev)
    scores = torch.empty((B, L, k), dtype=torch.float32, device=dev)
    index = torch.empty((B, L, k), dtype=torch.int64, device=dev)
    counts = torch.empty((B, L), dtype=torch.int32, device=dev)
    lib = _lib.load()
    with _dev(dev):
        nbytes = lib.lcr_rpn_select_workspace_bytes(B, L, k)
        ws = _workspace(nbytes, dev)
        check(lib.lcr_rpn_select_f32(levels, L, B, C.byref(cfg), boxes.data_ptr(), scores.data_ptr(), index.data_ptr(),
                                     counts.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "rpn_select")
    return boxes, scores, index, counts


def nms_batched(boxes: torch.Tensor, scores: Optional[torch.Tensor], iou_threshold: float, *, post_n: int,
                counts: Optional[torch.Tensor] = None, score_thresh: Optional[float] = None,
                category: Optional[torch.Tensor] = None, cpu_threshold: bool = False):
    """a8 — greedy NMS over S segments: boxes [S, stride, 4]; scores [S, stride] or None (already
    sorted); counts [S] i32 or None.  Returns keep [S, post_n] i64 (original in-segment indices, score
    order) and keep_counts [S] i32.

    Threshold semantics (they differ only for a pair whose fp32 IoU equals float32(thr) exactly): torchvision's CUDA op
    — what the reference executes on a GPU — compares the fp32 IoU with the threshold ROUNDED TO fp32; its CPU op
    compares with the double.  The C-ABI takes a double and implements the CPU rule, so the CUDA rule is obtained by
    pre-rounding the threshold here (default); cpu_threshold=True passes the double through (the CPU-generated golden
    vectors and the oracle follow that rule)."""
    if not cpu_threshold:
        iou_threshold = _round_f32(iou_threshold)
    _need_cuda(boxes, scores, counts, category)
    b = _f32c(boxes)
    S, stride = b.shape[0], b.shape[1]
    s = None if scores is None else _f32c(scores)
    cn = None if counts is None else counts.to(torch.int32).contiguous()
    cat = None if category is None else category.to(torch.int32).contiguous()
    keep = torch.empty((S, post_n), dtype=torch.int64, device=b.device)
    kc = torch.zeros((S,), dtype=torch.int32, device=b.device)
    if S == 0 or stride == 0:
        return keep, kc
    lib = _lib.load()
    with _dev(b.device):
        nbytes = lib.lcr_nms_workspace_bytes(S, stride)
        ws = _workspace(nbytes, b.device)
        check(lib.lcr_nms_f32(b.data_ptr(), _ptr(s), _ptr(cat), _ptr(cn), S, stride, float(iou_threshold),
                              0.0 if score_thresh is None else float(score_thresh), 0 if score_thresh is None else 1,
                              int(post_n), keep.data_ptr(), kc.data_ptr(), ws.data_ptr(), ws.numel(), _stream()), "nms")
    return keep, kc


def gather_kept(boxes: torch.Tensor, scores: Optional[torch.Tensor], keep: torch.Tensor, keep_counts: torch.Tensor,
                want_rois: bool = True, want_valid: bool = False):
    """proposals[keep] — returns (boxes [S,post_n,4], scores [S,post_n] | None, rois [S*post_n,5] | None
    [, valid [S*post_n] u8]); padding rows are zero boxes with roi batch index -1 / valid 0."""
    _need_cuda(boxes, keep, keep_counts)
    b = _f32c(boxes)
    S, in_stride = b.shape[0], b.shape[1]
    post_n = keep.shape[1]
    s = None if scores is None else _f32c(scores)
    ob = torch.empty((S, post_n, 4), dtype=torch.float32, device=b.device)
    osc = None if s is None else torch.empty((S, post_n), dtype=torch.float32, device=b.device)
    rois = torch.empty((S * post_n, 5), dtype=torch.float32, device=b.device) if want_rois else None
    valid = torch.empty((S * post_n,), dtype=torch.uint8, device=b.device) if want_valid else None
    if S:
        with _dev(b.device):
            check(_lib.load().lcr_gather_kept_f32(b.data_ptr(), _ptr(s), keep.data_ptr(), keep_counts.data_ptr(), S, in_stride,
                                                  post_n, ob.data_ptr(), _ptr(osc), _ptr(rois), _ptr(valid), _stream()),
                  "gather_kept")
    if want_valid:
        return ob, osc, rois, valid
    return ob, osc, rois


def level_map(boxes: torch.Tensor, k_min: int = 2, k_max: int = 5, canonical_scale: float = 224.0, canonical_level: int = 4,
              eps: float = 1e-6) -> torch.Tensor:
    """a11 — LevelMapper (TV:ops/poolers.py:73-84).  boxes [K,4] or rois [K,5] -> levels [K] i32."""
    _need_cuda(boxes)
    b = _f32c(boxes)
    K, bs = b.shape
    lv = torch.empty((K,), dtype=torch.int32, device=b.device)
    if K:
        with _dev(b.device):
            check(_lib.load().lcr_level_map_f32(b.data_ptr(), bs, K, k_min, k_max, float(canonical_scale), canonical_level,
                                                float(eps), lv.data_ptr(), _stream()), "level_map")
    return lv


def to_nhwc(x: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """[N,C,H,W] tensor -> same logical tensor in channels_last memory, through the tiled transpose
    kernel when a copy is needed (no-op for tensors that already are NHWC-dense and no `out`).
    `out`: optional caller-owned channels_last fp32 tensor of the same logical shape (a serving loop reuses it)."""
    _need_cuda(x, out)
    N, Cc, H, W = x.shape
    nhwc_strides = (H * W * Cc, 1, W * Cc, Cc)
    if out is not None and (tuple(out.shape) != (N, Cc, H, W) or out.stride() != nhwc_strides or out.dtype != torch.float32):
        raise _lib.LcrError("to_nhwc: out must be a channels_last fp32 tensor of the input's logical shape")
    if x.dtype == torch.float32 and x.stride() == nhwc_strides:
        if out is None:
            return x
        out.copy_(x)
        return out
    xc = _f32c(x)
    if out is None:
        out = torch.empty_strided((N, Cc, H, W), nhwc_strides, dtype=torch.float32, device=x.device)
    with _dev(x.device):
        check(_lib.load().lcr_nchw_to_nhwc_f32(xc.data_ptr(), out.data_ptr(), N, Cc, H, W, _stream()), "nchw_to_nhwc")
    return out


def to_nchw(x: torch.Tensor) -> torch.Tensor:
    """Inverse of to_nhwc: channels_last memory -> contiguous [N,C,H,W]."""
    _need_cuda(x)
    N, Cc, H, W = x.shape
    if x.is_contiguous():
        return x
    if x.stride() != (H * W * Cc, 1, W * Cc, Cc):
        return x.contiguous()
    out = torch.empty((N, Cc, H, W), dtype=torch.float32, device=x.device)
    with _dev(x.device):
        check(_lib.load().lcr_nhwc_to_nchw_f32(x.data_ptr(), out.data_ptr(), N, Cc, H, W, _stream()), "nhwc_to_nchw")
    return out


def _feat_levels(feats: Sequence[torch.Tensor], scales: Sequence[float]):
    L = len(feats)
    arr = (LcrFeatLevel * L)()
    for l, (f, sc) in enumerate(zip(feats, scales)):
        N, Cc, H, W = f.shape
        sn, s_c, sh, sw = f.stride()
        lv = arr[l]
        lv.data, lv.N, lv.H, lv.W = f.data_ptr(), N, H, W
        lv.sn, lv.sc, lv.sh, lv.sw = sn, s_c, sh, sw
        lv.spatial_scale = float(sc)
    return arr


def _dense(f: torch.Tensor) -> bool:
    N, Cc, H, W = f.shape
    return f.is_contiguous() or f.stride() == (H * W * Cc, 1, W * Cc, Cc)


def roi_align_fwd(feats: Sequence[torch.Tensor], scales: Sequence[float], rois: torch.Tensor,
                  roi_level: Optional[torch.Tensor], output_size, sampling_ratio: int, aligned: bool,
                  out: Optional[torch.Tensor] = None, cpu_coords: Optional[bool] = None) -> torch.Tensor:
    """a9/a11 — RoIAlign forward over one or several levels.  feats: logical [N,C,H,W] fp32 tensors
    (NCHW-contiguous or channels_last; other stridings are copied).  rois [K,5].  `out`: optional
    caller-owned contiguous fp32 buffer of at least K*C*PH*PW elements (a serving loop reuses it)."""
    feats = [f if (f.dtype == torch.float32 and _dense(f)) else _f32c(f) for f in feats]
    r = _f32c(rois).reshape(-1, 5)
    _need_cuda(r, roi_level, *feats)
    PH, PW = (output_size, output_size) if isinstance(output_size, int) else output_size
    K, Cc = r.shape[0], feats[0].shape[1]
    if out is None:
        out = torch.empty((K, Cc, PH, PW), dtype=torch.float32, device=r.device)
    else:
        if out.dtype != torch.float32 or not out.is_contiguous() or out.numel() < K * Cc * PH * PW or out.device != r.device:
            raise _lib.LcrError("roi_align_fwd: out must be a contiguous fp32 CUDA buffer of at least K*C*PH*PW elements")
        out = out.reshape(-1)[: K * Cc * PH * PW].view(K, Cc, PH, PW)
    if K == 0:
        return out
    lvl = None if roi_level is None else roi_level.to(torch.int32).contiguous()
    with _dev(r.device):
        check(_lib.load().lcr_roi_align_fwd_f32(_feat_levels(feats, scales), len(feats), Cc, r.data_ptr(), _ptr(lvl), K, PH, PW,
                                                int(sampling_ratio), roi_align_flags(aligned, cpu_coords), out.data_ptr(),
                                                _stream()),
              "roi_align_fwd")
    return out


def roi_align_bwd(grad_out: torch.Tensor, grads: Sequence[torch.Tensor], scales: Sequence[float], rois: torch.Tensor,
                  roi_level: Optional[torch.Tensor], sampling_ratio: int, aligned: bool, zero_grad: bool = True,
                  cpu_coords: Optional[bool] = None) -> None:
    """a10 — RoIAlign backward: accumulates into `grads` (dense NCHW or channels_last tensors of the
    forward feature shapes), zero-filling them first when zero_grad."""
    g = _f32c(grad_out)
    r = _f32c(rois).reshape(-1, 5)
    _need_cuda(g, r, roi_level, *grads)
    for t in grads:
        if t.dtype != torch.float32 or not _dense(t):
            raise _lib.LcrError("roi_align_bwd: grad buffers must be dense fp32 (NCHW or channels_last)")
    K, Cc, PH, PW = g.shape
    lvl = None if roi_level is None else roi_level.to(torch.int32).contiguous()
    with _dev(g.device):
        check(_lib.load().lcr_roi_align_bwd_f32(g.data_ptr(), _feat_levels(grads, scales), len(grads), Cc, r.data_ptr(), _ptr(lvl),
                                                K, PH, PW, int(sampling_ratio), roi_align_flags(aligned, cpu_coords),
                                                1 if zero_grad else 0, _stream()), "roi_align_bwd")


def paste_masks(probs: torch.Tensor, boxes: torch.Tensor, img_h: int, img_w: int, threshold: float = 0.5,
                on_value: int = 255, valid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a13 — batched paste/threshold (src/custom_maskrcnn.py:276-295, src/utils/mask_utils.py:149-171).
    probs [N,M,M] f32, boxes [N,4] f32 -> uint8 [N,H,W] in {0,on_value}."""
    _need_cuda(probs, boxes, valid, out)
    p, b = _f32c(probs), _f32c(boxes).reshape(-1, 4)
    N = b.shape[0]
    M = p.shape[-1] if p.numel() else 28
    if out is None:
        out = torch.empty((N, img_h, img_w), dtype=torch.uint8, device=b.device)
    elif out.dtype != torch.uint8 or not out.is_contiguous() or out.numel() < N * img_h * img_w:
        raise _lib.LcrError("paste_masks: out must be a contiguous uint8 buffer of at least N*H*W bytes")
    v = None if valid is None else valid.to(torch.uint8).contiguous()
    if N:
        with _dev(b.device):
            check(_lib.load().lcr_paste_masks_u8(p.data_ptr(), b.data_ptr(), _ptr(v), N, M, img_h, img_w, float(threshold),
                                                 int(on_value), out.data_ptr(), _stream()), "paste_masks")
    return out


def paste_masks_tv(probs: torch.Tensor, boxes: torch.Tensor, img_h: int, img_w: int, padding: int = 1,
                   valid: Optional[torch.Tensor] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """a13, torchvision variant — torchvision.models.detection.roi_heads.paste_masks_in_image (TV:models/detection/
    roi_heads.py:405-501): probs [N,M,M] (or [N,1,M,M]) f32, boxes [N,4] -> float32 [N,1,H,W] probabilities pasted into
    the box expanded by (M + 2 padding)/M, zero elsewhere."""
    _need_cuda(probs, boxes, valid, out)
    p, b = _f32c(probs), _f32c(boxes).reshape(-1, 4)
    N = b.shape[0]
    M = p.shape[-1] if p.numel() else 28
    if out is None:
        out = torch.empty((N, 1, img_h, img_w), dtype=torch.float32, device=b.device)
    elif out.dtype != torch.float32 or not out.is_contiguous() or out.numel() < N * img_h * img_w:
        raise _lib.LcrError("paste_masks_tv: out must be a contiguous float32 buffer of at least N*H*W elements")
    v = None if valid is None else valid.to(torch.uint8).contiguous()
    if N:
        with _dev(b.device):
            check(_lib.load().lcr_paste_masks_tv_f32(p.data_ptr(), b.data_ptr(), _ptr(v), N, M, img_h, img_w, int(padding),
                                                     out.data_ptr(), _stream()), "paste_masks_tv")
    return out


def rpn_concat_levels(boxes: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor, coordinate_trick: bool):
    """a12 — the cross-level step of torchvision's RPN (see lcr_rpn_concat_levels_f32): boxes [B,L,k,4], scores [B,L,k],
    counts [B,L] -> (cat_boxes [B,L*k,4], nms_boxes [B,L*k,4], cat_scores [B,L*k], cat_level [B,L*k] i32, cat_counts [B])."""
    _need_cuda(boxes, scores, counts)
    b, s = _f32c(boxes), _f32c(scores)
    B, L, k = s.shape
    cn = counts.to(torch.int32).contiguous()
    cb = torch.empty((B, L * k, 4), dtype=torch.float32, device=b.device)
    nb = torch.empty((B, L * k, 4), dtype=torch.float32, device=b.device)
    cs = torch.empty((B, L * k), dtype=torch.float32, device=b.device)
    cl = torch.empty((B, L * k), dtype=torch.int32, device=b.device)
    cc = torch.empty((B,), dtype=torch.int32, device=b.device)
    if B:
        with _dev(b.device):
            check(_lib.load().lcr_rpn_concat_levels_f32(b.data_ptr(), s.data_ptr(), cn.data_ptr(), B, L, k, 1 if coordinate_trick else 0,
                                                        cb.data_ptr(), nb.data_ptr(), cs.data_ptr(), cl.data_ptr(), cc.data_ptr(),
                                                        _stream()), "rpn_concat_levels")
    return cb, nb, cs, cl, cc


def pack_records(boxes: torch.Tensor, scores: torch.Tensor, counts: torch.Tensor) -> torch.Tensor:
    """Detection records [S, stride, 6] = (x1,y1,x2,y2,score,label=1), zero padded."""
    _need_cuda(boxes, scores, counts)
    b, s = _f32c(boxes), _f32c(scores)
    S, stride = s.shape
    cn = counts.to(torch.int32).contiguous()
    rec = torch.empty((S, stride, 6), dtype=torch.float32, device=b.device)
    if S:
        with _dev(b.device):
            check(_lib.load().lcr_pack_records_f32(b.data_ptr(), s.data_ptr(), cn.data_ptr(), S, stride, rec.data_ptr(), _stream()),
                  "pack_records")
    return rec


# ------------------------------------------------------------------------------------------------
# training-side siblings
def box_iou(boxes: torch.Tensor, gt: torch.Tensor) -> torch.Tensor:
    """torchvision.ops.box_iou (TV:ops/boxes.py:308-370): [N,4] x [G,4] -> [N,G], one kernel."""
    _need_cuda(boxes, gt)
    a, b = _f32c(boxes).reshape(-1, 4), _f32c(gt).reshape(-1, 4)
    N, G = a.shape[0], b.shape[0]
    out = torch.empty((N, G), dtype=torch.float32, device=a.device)
    if N and G:
        with _dev(a.device):
            check(_lib.load().lcr_box_iou_f32(a.data_ptr(), N, b.data_ptr(), G, out.data_ptr(), _stream()), "box_iou")
    return out


def box_iou_max(boxes: torch.Tensor, gt: torch.Tensor):
    """Fused `box_iou(boxes, gt).max(dim=1)` (src/components/rpn.py:72-73, src/custom_maskrcnn.py:221-222):
    (max_iou [N] f32, argmax [N] i64) without the [N,G] matrix."""
    _need_cuda(boxes, gt)
    a, b = _f32c(boxes).reshape(-1, 4), _f32c(gt).reshape(-1, 4)
    N, G = a.shape[0], b.shape[0]
    if G == 0:
        raise _lib.LcrError("box_iou_max: no ground-truth boxes (torch.max over an empty dimension is an error too)")
    mx = torch.empty((N,), dtype=torch.float32, device=a.device)
    am = torch.empty((N,), dtype=torch.int64, device=a.device)
    if N:
        with _dev(a.device):
            check(_lib.load().lcr_box_iou_max_f32(a.data_ptr(), N, b.data_ptr(), G, mx.data_ptr(), am.data_ptr(), _stream()),
                  "box_iou_max")
    return mx, am


def match_boxes(boxes: torch.Tensor, gt: torch.Tensor, pos_thr: float, neg_thr: Optional[float] = None):
    """Fused matcher — `ious = box_iou(boxes, gt); max_iou, idx = ious.max(dim=1); pos = max_iou >= pos_thr;
    neg = max_iou < neg_thr` plus the two `.sum()`s (src/components/rpn.py:72-81 with 0.5 / 0.3;
    src/custom_maskrcnn.py:221-225, 249-251 with 0.4) in one kernel.  Returns (max_iou [N] f32, argmax [N] i64,
    pos_mask [N] bool, neg_mask [N] bool, counts [2] i32 on the device: number of positives, number of negatives).
    neg_thr=None: negatives are the complement threshold (`max_iou < pos_thr`), the box head's background."""
    _need_cuda(boxes, gt)
    a, b = _f32c(boxes).reshape(-1, 4), _f32c(gt).reshape(-1, 4)
    N, G = a.shape[0], b.shape[0]
    if G == 0:
        raise _lib.LcrError("match_boxes: no ground-truth boxes (the reference returns its no-target loss before matching)")
    if neg_thr is None:
        neg_thr = pos_thr
    mx = torch.empty((N,), dtype=torch.float32, device=a.device)
    am = torch.empty((N,), dtype=torch.int64, device=a.device)
    pos = torch.empty((N,), dtype=torch.uint8, device=a.device)
    neg = torch.empty((N,), dtype=torch.uint8, device=a.device)
    counts = torch.empty((2,), dtype=torch.int32, device=a.device)
    with _dev(a.device):
        check(_lib.load().lcr_match_boxes_f32(a.data_ptr(), N, b.data_ptr(), G, float(pos_thr), float(neg_thr), mx.data_ptr(),
                                              am.data_ptr(), pos.data_ptr(), neg.data_ptr(), counts.data_ptr(), _stream()),
              "match_boxes")
    return mx, am, pos.view(torch.bool), neg.view(torch.bool), counts


def mask_targets(gt_masks: torch.Tensor, boxes: torch.Tensor, gt_index: Optional[torch.Tensor] = None, mask_size: int = 28) -> torch.Tensor:
    """Batched extract_mask_target (src/utils/mask_utils.py:6-46): gt_masks [G,H,W] uint8, boxes [K,4],
    gt_index [K] i64 (None: mask k) -> [K,M,M] f32 bilinear crops."""
    _need_cuda(gt_masks, boxes, gt_index)
    m = gt_masks if gt_masks.dtype == torch.uint8 else gt_masks.to(torch.uint8)
    m = m.contiguous()
    b = _f32c(boxes).reshape(-1, 4)
    K = b.shape[0]
    G, H, W = m.shape
    idx = None if gt_index is None else gt_index.to(torch.int64).contiguous()
    out = torch.empty((K, mask_size, mask_size), dtype=torch.float32, device=b.device)
    if K:
        if G == 0:
            raise _lib.LcrError("mask_targets: no ground-truth masks")
        with _dev(b.device):
            check(_lib.load().lcr_mask_targets_f32(m.data_ptr(), G, H, W, b.data_ptr(), _ptr(idx), K, int(mask_size), out.data_ptr(),
                                                   _stream()), "mask_targets")
    return out


def mask_tail(mask_logits: torch.Tensor, mask_size: int = 28, cls: int = 1) -> torch.Tensor:
    """Fused tail of the mask head: bilinear (align_corners=False) resize of class `cls` of
    mask_logits [K,num_classes,m,m] to mask_size + sigmoid -> probs [K,mask_size,mask_size]
    (src/components/mask_head.py:52-58 + src/custom_maskrcnn.py:273-274)."""
    _need_cuda(mask_logits)
    x = _f32c(mask_logits)
    K, Cc, m, m2 = x.shape
    if m != m2:
        raise _lib.LcrError("mask_tail: square mask logits expected")
    out = torch.empty((K, mask_size, mask_size), dtype=torch.float32, device=x.device)
    if K:
        with _dev(x.device):
            check(_lib.load().lcr_mask_tail_f32(x.data_ptr(), K, Cc, int(cls), m, int(mask_size), out.data_ptr(), _stream()), "mask_tail")
    return out


def mask_region_counts(masks: torch.Tensor, rects: torch.Tensor, rect_offsets: torch.Tensor, boxes: Optional[torch.Tensor] = None,
                       threshold: int = 0):
    """Per-detection pixel counts for tile stitching (src/visualize.py:106-130): masks [N,H,W] uint8,
    rects [R,4] int32 (x0,y0,x1,y1 half-open, clipped), rect_offsets [N+1] int32 -> (total [N] i32, in_region [R] i32)."""
    _need_cuda(masks, rects, rect_offsets, boxes)
    if masks.dtype != torch.uint8:
        raise _lib.LcrError("mask_region_counts: uint8 masks expected")
    m = masks.contiguous()
    N, H, W = m.shape
    r = rects.to(torch.int32).contiguous().reshape(-1, 4)
    ro = rect_offsets.to(torch.int32).contiguous()
    b = None if boxes is None else _f32c(boxes).reshape(-1, 4)
    total = torch.zeros((N,), dtype=torch.int32, device=m.device)
    inreg = torch.zeros((max(r.shape[0], 1),), dtype=torch.int32, device=m.device)
    if N:
        with _dev(m.device):
            check(_lib.load().lcr_mask_region_counts_u8(m.data_ptr(), N, H, W, _ptr(b), r.data_ptr() if r.numel() else None, ro.data_ptr(),
                                                        int(threshold), total.data_ptr(), inreg.data_ptr(), _stream()),
                  "mask_region_counts")
    return total, inreg[: r.shape[0]]

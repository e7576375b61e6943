"""numpy/ctypes front-end of the CPU oracle (oracle/lcr_oracle.c).

TEST INFRASTRUCTURE ONLY — see the header of lcr_oracle.c.  Importable from tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs; never from the product
package.  Every function takes and returns numpy arrays on the host.
"""
from __future__ import annotations

import ctypes as C
import math
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "liblcr_oracle.so")
_lib = None

XFORM_CLIP = math.log(1000.0 / 16)


def build(force: bool = False) -> str:
    """Compile oracle/lcr_oracle.c with gcc (see oracle/Makefile)."""
    src = os.path.join(_HERE, "lcr_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s", "-B"], check=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        _lib.orc_num_threads.restype = C.c_int
        _lib.orc_rpn_select_segment.restype = C.c_int
        _lib.orc_nms.restype = C.c_int
    return _lib


def num_threads() -> int:
    return int(lib().orc_num_threads())


def usable_cores() -> int:
    """Cores this process may run on (cgroup/affinity aware), not what OMP_NUM_THREADS says."""
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except AttributeError:  # pragma: no cover
        return max(1, os.cpu_count() or 1)


def set_num_threads(n: int = 0) -> int:
    """Use n OpenMP threads (0 = every usable core) regardless of an inherited OMP_NUM_THREADS (torchrun exports
    OMP_NUM_THREADS=1 to its ranks, which made the r01 multi-GPU reference arm single-threaded).  Returns the count."""
    n = int(n) if n and n > 0 else usable_cores()
    lib().orc_set_num_threads(C.c_int(n))
    return num_threads()


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, ty=C.c_float):
    return None if a is None else a.ctypes.data_as(C.POINTER(ty))


def base_anchors(sizes=(32, 64, 128), aspect_ratios=(0.5, 1.0, 2.0)) -> np.ndarray:
    """Base anchors of src/components/anchor_generator.py:15-27: float64 formula, then fp32."""
    rows = []
    for size in sizes:
        for ratio in aspect_ratios:
            area = size * size
            h = math.sqrt(area / ratio)
            w = h * ratio
            rows.append([-w / 2, -h / 2, w / 2, h / 2])
    return np.asarray(rows, dtype=np.float64).astype(np.float32)


def anchors(h, w, stride, base) -> np.ndarray:
    base = _f32(base)
    A = base.shape[0]
    out = np.empty((h * w * A, 4), np.float32)
    lib().orc_anchors(_p(out), C.c_int(h), C.c_int(w), C.c_int(stride), _p(base), C.c_int(A))
    return out


def clip_boxes(boxes, img_h, img_w) -> np.ndarray:
    b = _f32(boxes).copy()
    lib().orc_clip_boxes(_p(b), C.c_int(b.shape[0]), C.c_float(img_h), C.c_float(img_w))
    return b


def filter_small_boxes(boxes, min_size) -> np.ndarray:
    b = _f32(boxes)
    keep = np.zeros(b.shape[0], np.uint8)
    lib().orc_filter_small_boxes(_p(b), C.c_int(b.shape[0]), C.c_float(min_size), _p(keep, C.c_uint8))
    return keep.astype(bool)


def box_decode(deltas, anc, weights=(1.0, 1.0, 1.0, 1.0), xform_clip=XFORM_CLIP, img_h=0.0, img_w=0.0):
    d, a = _f32(deltas), _f32(anc)
    w = _f32(np.asarray(weights))
    out = np.empty_like(a)
    lib().orc_box_decode(_p(d), _p(a), C.c_int(a.shape[0]), _p(w), C.c_float(xform_clip),
                         C.c_float(img_h), C.c_float(img_w), _p(out))
    return out


def rpn_select(obj, *, anchors=None, base=None, stride=4, k=250, score_thresh=0.3, score_strict=True,
               min_size=10.0, img_h=520, img_w=704, topk_on_sigmoid=True, deltas=None,
               decode_weights=(1.0, 1.0, 1.0, 1.0), xform_clip=XFORM_CLIP):
    """One (image, level) segment.  obj [A,h,w].  Returns (boxes [n,4], scores [n], index [n] i64)."""
    obj = _f32(obj)
    A, h, w = obj.shape
    anc = None if anchors is None else _f32(anchors)
    bs = None if base is None else _f32(base)
    dl = None if deltas is None else _f32(deltas)
    wt = _f32(np.asarray(decode_weights))
    boxes = np.zeros((k, 4), np.float32)
    scores = np.zeros((k,), np.float32)
    index = np.zeros((k,), np.int64)
    n = lib().orc_rpn_select_segment(
        _p(obj), _p(dl), _p(anc), _p(bs), C.c_int(A), C.c_int(h), C.c_int(w), C.c_int(stride), C.c_int(k),
        C.c_float(score_thresh), C.c_int(1 if score_strict else 0), C.c_float(min_size), C.c_int(img_h),
        C.c_int(img_w), C.c_int(1 if topk_on_sigmoid else 0), _p(wt), C.c_float(xform_clip), _p(boxes),
        _p(scores), _p(index, C.c_int64))
    return boxes[:n].copy(), scores[:n].copy(), index[:n].copy()


def rpn_select_batch(obj, *, base, stride=4, k=250, score_thresh=0.3, score_strict=True, min_size=10.0,
                     img_h=520, img_w=704, topk_on_sigmoid=True, anchors=None):
    """Batched single-level driver (threads over images).  obj [B,A,h,w] -> padded outputs + counts."""
    obj = _f32(obj)
    B, A, h, w = obj.shape
    bs = _f32(base)
    anc = None if anchors is None else _f32(anchors)
    boxes = np.zeros((B, k, 4), np.float32)
    scores = np.zeros((B, k), np.float32)
    index = np.zeros((B, k), np.int64)
    counts = np.zeros((B,), np.int32)
    lib().orc_rpn_select_batch(
        _p(obj), _p(anc), _p(bs), C.c_int(B), C.c_int(A), C.c_int(h), C.c_int(w), C.c_int(stride), C.c_int(k),
        C.c_float(score_thresh), C.c_int(1 if score_strict else 0), C.c_float(min_size), C.c_int(img_h),
        C.c_int(img_w), C.c_int(1 if topk_on_sigmoid else 0), _p(boxes), _p(scores), _p(index, C.c_int64),
        _p(counts, C.c_int))
    return boxes, scores, index, counts


def nms(boxes, scores=None, iou_thr=0.5, *, category=None, score_thresh=0.0, use_score_thresh=False,
        post_n=None) -> np.ndarray:
    b = _f32(boxes).reshape(-1, 4)
    n = b.shape[0]
    s = None if scores is None else _f32(scores)
    cat = None if category is None else np.ascontiguousarray(category, dtype=np.int32)
    post_n = n if post_n is None else post_n
    keep = np.zeros((max(post_n, 1),), np.int64)
    cnt = lib().orc_nms(_p(b), _p(s), _p(cat, C.c_int), C.c_int(n), C.c_double(iou_thr), C.c_float(score_thresh),
                        C.c_int(1 if use_score_thresh else 0), C.c_int(post_n), _p(keep, C.c_int64))
    return keep[:cnt].copy()


def nms_batch(boxes, scores, counts, iou_thr, *, score_thresh=0.0, use_score_thresh=False, post_n):
    b = _f32(boxes)
    S, stride = b.shape[0], b.shape[1]
    s = None if scores is None else _f32(scores)
    cn = None if counts is None else np.ascontiguousarray(counts, dtype=np.int32)
    keep = np.zeros((S, post_n), np.int64)
    kc = np.zeros((S,), np.int32)
    lib().orc_nms_batch(_p(b), _p(s), _p(cn, C.c_int), C.c_int(S), C.c_int(stride), C.c_double(iou_thr),
                        C.c_float(score_thresh), C.c_int(1 if use_score_thresh else 0), C.c_int(post_n),
                        _p(keep, C.c_int64), _p(kc, C.c_int))
    return keep, kc


def level_map(boxes, k_min=2, k_max=5, canonical_scale=224.0, canonical_level=4, eps=1e-6) -> np.ndarray:
    b = _f32(boxes)
    K, bs = b.shape
    lv = np.zeros((K,), np.int32)
    lib().orc_level_map(_p(b), C.c_int(bs), C.c_int(K), C.c_int(k_min), C.c_int(k_max), C.c_float(canonical_scale),
                        C.c_int(canonical_level), C.c_float(eps), _p(lv, C.c_int))
    return lv


def _strides(feat_shape, nhwc: bool):
    N, Cc, H, W = feat_shape
    if nhwc:
        return (H * W * Cc, 1, W * Cc, Cc)
    return (Cc * H * W, H * W, W, 1)


def roi_align_fwd(feat, rois, PH=7, PW=7, scale=0.25, sampling_ratio=2, aligned=False, cuda_coords=False) -> np.ndarray:
    """feat: logical [N,C,H,W] numpy array (any strides are honoured through a contiguous copy).
    cuda_coords: round the sample coordinates as torchvision's CUDA op does (FMA-contracted) instead of its CPU op."""
    f = _f32(feat)
    N, Cc, H, W = f.shape
    r = _f32(rois).reshape(-1, 5)
    K = r.shape[0]
    out = np.zeros((K, Cc, PH, PW), np.float32)
    sn, sc, sh, sw = _strides(f.shape, False)
    lib().orc_roi_align_fwd(_p(f), C.c_int(N), C.c_int(Cc), C.c_int(H), C.c_int(W), C.c_int64(sn), C.c_int64(sc),
                            C.c_int64(sh), C.c_int64(sw), _p(r), C.c_int(K), C.c_int(PH), C.c_int(PW),
                            C.c_float(scale), C.c_int(sampling_ratio), C.c_int((1 if aligned else 0) | (2 if cuda_coords else 0)),
                            _p(out))
    return out


def roi_align_bwd(grad_out, rois, feat_shape, scale=0.25, sampling_ratio=2, aligned=False, cuda_coords=False) -> np.ndarray:
    g = _f32(grad_out)
    K, Cc, PH, PW = g.shape
    N, C2, H, W = feat_shape
    assert C2 == Cc
    r = _f32(rois).reshape(-1, 5)
    gi = np.zeros((N, Cc, H, W), np.float32)
    sn, sc, sh, sw = _strides(gi.shape, False)
    lib().orc_roi_align_bwd(_p(g), C.c_int(N), C.c_int(Cc), C.c_int(H), C.c_int(W), C.c_int64(sn), C.c_int64(sc),
                            C.c_int64(sh), C.c_int64(sw), _p(r), C.c_int(K), C.c_int(PH), C.c_int(PW),
                            C.c_float(scale), C.c_int(sampling_ratio), C.c_int((1 if aligned else 0) | (2 if cuda_coords else 0)),
                            _p(gi))
    return gi


def multiscale_roi_align_fwd(feats, scales, rois, levels, PH=7, PW=7, sampling_ratio=2, aligned=False, cuda_coords=False):
    """MultiScaleRoIAlign (TV:ops/poolers.py:147-227): per-level roi_align scattered to roi order."""
    r = _f32(rois).reshape(-1, 5)
    Cc = feats[0].shape[1]
    out = np.zeros((r.shape[0], Cc, PH, PW), np.float32)
    for l, (f, s) in enumerate(zip(feats, scales)):
        sel = np.nonzero(np.asarray(levels) == l)[0]
        if sel.size:
            out[sel] = roi_align_fwd(f, r[sel], PH, PW, s, sampling_ratio, aligned, cuda_coords)
    return out


def paste_masks(probs, boxes, H, W, thr=0.5, on_value=255, valid=None, out=None) -> np.ndarray:
    p = _f32(probs)
    N, M = p.shape[0], p.shape[-1]
    b = _f32(boxes).reshape(-1, 4)
    v = None if valid is None else np.ascontiguousarray(valid, dtype=np.uint8)
    if out is None:
        out = np.zeros((N, H, W), np.uint8)
    lib().orc_paste_masks(_p(p), _p(b), _p(v, C.c_uint8), C.c_int(N), C.c_int(M), C.c_int(H), C.c_int(W),
                          C.c_float(thr), C.c_uint8(on_value), _p(out, C.c_uint8))
    return out


def paste_masks_tv(probs, boxes, H, W, padding=1) -> np.ndarray:
    """torchvision's paste_masks_in_image (TV:models/detection/roi_heads.py:405-501) -> float32 [N,H,W] probabilities."""
    p = _f32(probs)
    N, M = p.shape[0], p.shape[-1]
    b = _f32(boxes).reshape(-1, 4)
    out = np.zeros((N, H, W), np.float32)
    lib().orc_paste_masks_tv(_p(p), _p(b), C.c_int(N), C.c_int(M), C.c_int(H), C.c_int(W), C.c_int(padding), _p(out))
    return out


def tv_base_anchors(sizes, aspect_ratios) -> np.ndarray:
    """torchvision AnchorGenerator.generate_anchors (TV:models/detection/anchor_utils.py:58-78): ratio = h/w,
    h_ratios = sqrt(ratios), w_ratios = 1/h_ratios, all in fp32, base = round([-w,-h,w,h]/2) (half to even)."""
    scales = np.asarray(sizes, np.float32)
    ratios = np.asarray(aspect_ratios, np.float32)
    h_r = np.sqrt(ratios)
    w_r = (np.float32(1.0) / h_r).astype(np.float32)
    ws = (w_r[:, None] * scales[None, :]).reshape(-1)
    hs = (h_r[:, None] * scales[None, :]).reshape(-1)
    base = (np.stack([-ws, -hs, ws, hs], axis=1) / np.float32(2.0)).astype(np.float32)
    return np.round(base).astype(np.float32)


def tv_rpn_filter_proposals(objs, deltas, bases, strides, img_h, img_w, *, k, post_n, nms_thresh, score_thresh=0.0,
                            min_size=1e-3, coordinate_trick=None):
    """torchvision RegionProposalNetwork.forward's inference path for ONE image (TV:models/detection/rpn.py:231-297,
    :339-367): per level top-k on the LOGITS, decode (BoxCoder weights 1,1,1,1), sigmoid, clip, remove_small_boxes(min_size),
    `score >= score_thresh`, batched_nms over the levels, first post_n.
    objs / deltas: per-level [A,h,w] / [4A,h,w]; bases: per-level [A,4] (tv_base_anchors).  coordinate_trick: None = what
    torchvision does on CPU (trick iff 4 * boxes <= 4000), True/False to force.
    Returns (boxes [n,4], scores [n], level [n], top_idx: per-level flat indices of the pre-NMS top-k)."""
    cb, cs, cl, top = [], [], [], []
    for l, (o, d) in enumerate(zip(objs, deltas)):
        n = int(np.prod(o.shape))
        kk = min(k, n)
        # the unfiltered top-k index list (what _get_top_n_idx returns)
        _, _, ti = rpn_select(o, base=bases[l], stride=strides[l], k=kk, score_thresh=-1.0, score_strict=False, min_size=-1.0,
                              img_h=img_h, img_w=img_w, topk_on_sigmoid=False, deltas=d)
        top.append(ti)
        b, s, _ = rpn_select(o, base=bases[l], stride=strides[l], k=kk, score_thresh=score_thresh, score_strict=False,
                             min_size=min_size, img_h=img_h, img_w=img_w, topk_on_sigmoid=False, deltas=d)
        cb.append(b)
        cs.append(s)
        cl.append(np.full((len(s),), l, np.int32))
    boxes, scores, level = np.concatenate(cb), np.concatenate(cs), np.concatenate(cl)
    if coordinate_trick is None:
        coordinate_trick = boxes.size <= 4000
    if len(boxes) == 0:
        return boxes, scores, level, top
    if coordinate_trick:          # _batched_nms_coordinate_trick: fp32 offsets, plain nms
        shift = np.float32(boxes.max()) + np.float32(1.0)
        off = (level.astype(np.float32) * shift).astype(np.float32)
        keep = nms((boxes + off[:, None]).astype(np.float32), scores, nms_thresh)
    else:                         # _batched_nms_vanilla: per-level nms, result sorted by score
        keep = nms(boxes, scores, nms_thresh, category=level)
    keep = keep[:post_n]
    return boxes[keep], scores[keep], level[keep], top


def pack_records(boxes, scores, counts) -> np.ndarray:
    b, s = _f32(boxes), _f32(scores)
    S, stride = s.shape
    cn = np.ascontiguousarray(counts, dtype=np.int32)
    rec = np.zeros((S, stride, 6), np.float32)
    lib().orc_pack_records(_p(b), _p(s), _p(cn, C.c_int), C.c_int(S), C.c_int(stride), _p(rec))
    return rec


def box_iou(boxes, gt, want_matrix=True):
    """torchvision.ops.box_iou + `.max(dim=1)`.  Returns (iou [N,G] or None, max [N], argmax [N])."""
    a, b = _f32(boxes).reshape(-1, 4), _f32(gt).reshape(-1, 4)
    N, G = a.shape[0], b.shape[0]
    iou = np.zeros((N, G), np.float32) if want_matrix else None
    mx, am = np.zeros((N,), np.float32), np.zeros((N,), np.int64)
    lib().orc_box_iou(_p(a), C.c_int(N), _p(b), C.c_int(G), _p(iou), _p(mx), _p(am, C.c_int64))
    return iou, mx, am


def match_boxes(boxes, gt, pos_thr, neg_thr=None):
    """The threshold masks and their sums that follow the row max: src/components/rpn.py:76-81 (0.5 / 0.3),
    src/custom_maskrcnn.py:224-225, 251 (0.4).  A float tensor compared with a Python scalar is compared in fp32 (ATen);
    NaN rows fall in neither mask.  Returns (max [N] f32, argmax [N] i64, pos [N] bool, neg [N] bool, counts [2] i32)."""
    _, mx, am = box_iou(boxes, gt, want_matrix=False)
    neg_thr = pos_thr if neg_thr is None else neg_thr
    with np.errstate(invalid="ignore"):
        pos = mx >= np.float32(pos_thr)
        neg = mx < np.float32(neg_thr)
    return mx, am, pos, neg, np.array([pos.sum(), neg.sum()], np.int32)


def mask_targets(gt_masks, boxes, gt_index=None, M=28) -> np.ndarray:
    """extract_mask_target (src/utils/mask_utils.py:6-46) for K (box, mask index) pairs."""
    m = np.ascontiguousarray(gt_masks, dtype=np.uint8)
    G, H, W = m.shape
    b = _f32(boxes).reshape(-1, 4)
    K = b.shape[0]
    idx = None if gt_index is None else np.ascontiguousarray(gt_index, dtype=np.int64)
    out = np.zeros((K, M, M), np.float32)
    lib().orc_mask_targets(_p(m, C.c_uint8), C.c_int(G), C.c_int(H), C.c_int(W), _p(b), _p(idx, C.c_int64), C.c_int(K), C.c_int(M),
                           _p(out))
    return out


def mask_tail(logits, M=28, cls=1) -> np.ndarray:
    """Bilinear resize of class `cls` of logits [K,Cc,m,m] to MxM + sigmoid."""
    x = _f32(logits)
    K, Cc, m, _ = x.shape
    out = np.zeros((K, M, M), np.float32)
    lib().orc_mask_tail(_p(x), C.c_int(K), C.c_int(Cc), C.c_int(cls), C.c_int(m), C.c_int(M), _p(out))
    return out


def mask_region_counts(masks, rects, rect_offsets, threshold=0):
    """Pixel counts per detection and per rectangle."""
    m = np.ascontiguousarray(masks, dtype=np.uint8)
    N, H, W = m.shape
    r = np.ascontiguousarray(rects, dtype=np.int32).reshape(-1, 4)
    ro = np.ascontiguousarray(rect_offsets, dtype=np.int32)
    total, inreg = np.zeros((N,), np.int32), np.zeros((max(len(r), 1),), np.int32)
    lib().orc_mask_region_counts(_p(m, C.c_uint8), C.c_int(N), C.c_int(H), C.c_int(W), _p(r, C.c_int), _p(ro, C.c_int),
                                 C.c_int(threshold), _p(total, C.c_int), _p(inreg, C.c_int))
    return total, inreg[: len(r)]

"""Seeded synthetic LIVECell-shaped inputs for the region pipeline.

numpy only (no torch import): the same arrays are produced in the authoring container (golden
fixtures, CPU oracle) and on the GPU box (parity tests, bench), so a seed identifies an input.

Geometry of the reference: 704x520 grayscale frames (src/preprocess_dataset.py:86-93), FPN level 0
at stride 4 -> 130x176 map, 9 anchors per location (src/components/anchor_generator.py:8-11).
"""
from __future__ import annotations

import numpy as np

IMG_H, IMG_W = 520, 704
STRIDE = 4
NUM_ANCHORS = 9


def _sigmoid64(x):
    return 1.0 / (1.0 + np.exp(-x.astype(np.float64)))


def make_objectness(B, A, h, w, n_cells, seed, k, anchor_choices=None):
    """RPN objectness logits [B, A, h, w]: N(-4,1) background plus +U(6,10) at n_cells random
    (a, y, x) "cell centre" anchors per image.  The top-(k+1) post-sigmoid fp32 values of every image
    are made unique with a >= 8-ulp gap (tie-free fixture: torch.topk's tie order is unspecified, so
    bit-exact index parity is only defined on tie-free data — SURVEY.md §7 "Tie semantics").
    anchor_choices: optional list of anchor ids the cells may use (default: all A)."""
    rng = np.random.RandomState(seed)
    out = np.empty((B, A, h, w), np.float32)
    for b in range(B):
        x = rng.normal(-4.0, 1.0, size=(A, h, w)).astype(np.float32)
        n = A * h * w
        nc = min(n_cells, n)
        if anchor_choices is None:
            flat = rng.choice(n, size=nc, replace=False)
        else:
            per = h * w
            pos = rng.choice(per, size=min(nc, per), replace=False)
            aa = rng.choice(np.asarray(anchor_choices), size=pos.size)
            flat = aa * per + pos
        xf = x.reshape(-1)
        xf[flat] += rng.uniform(6.0, 10.0, size=flat.size).astype(np.float32)
        _make_top_unique(xf, min(k + 1, n), rng)
        out[b] = xf.reshape(A, h, w)
    return out


def _make_top_unique(xf, m, rng, min_ulps=8, max_rounds=64):
    """Nudge logits in place until the top-m fp32 sigmoid values are pairwise >= min_ulps apart."""
    for _ in range(max_rounds):
        s = _sigmoid64(xf).astype(np.float32)
        top = np.argpartition(-s, m - 1)[:m] if m < s.size else np.arange(s.size)
        top = top[np.argsort(-s[top], kind="stable")]
        v = s[top].astype(np.float64)
        gap = v[:-1] - v[1:]
        need = min_ulps * np.spacing(s[top][:-1]).astype(np.float64)
        bad = np.nonzero(gap < need)[0]
        if bad.size == 0:
            return
        # move the lower element of each colliding pair down by a random, clearly resolvable amount
        xf[top[bad + 1]] -= rng.uniform(0.01, 0.05, size=bad.size).astype(np.float32)
    raise RuntimeError("could not make top-k unique")


def make_features(B, C, H, W, seed, nhwc=False):
    """FPN feature maps N(0,1).  Returns the logical [B,C,H,W] array; with nhwc=True the memory is
    [B,H,W,C] contiguous (a channels_last view)."""
    rng = np.random.RandomState(seed)
    if nhwc:
        f = rng.standard_normal((B, H, W, C)).astype(np.float32)
        return f.transpose(0, 3, 1, 2)
    return rng.standard_normal((B, C, H, W)).astype(np.float32)


def make_rois(K, seed, img_h=IMG_H, img_w=IMG_W, mode="anchor", batch=1, edge_cases=False):
    """RoIs [K,5] = (batch_idx, x1, y1, x2, y2).
    mode "anchor": anchor-shaped boxes (sizes 32/64/128 x ratios .5/1/2, clipped: 22..181 px), the
    reference's own proposal population; mode "fpn": sizes log-uniform 12..400 px, aspect
    exp(U(-.7,.7)).  edge_cases adds degenerate / out-of-range boxes up front."""
    rng = np.random.RandomState(seed)
    if mode == "anchor":
        size = rng.choice([32.0, 64.0, 128.0], size=K)
        ratio = rng.choice([0.5, 1.0, 2.0], size=K)
        hh = np.sqrt(size * size / ratio)
        ww = hh * ratio
    else:
        s = np.exp(rng.uniform(np.log(12.0), np.log(400.0), size=K))
        asp = np.exp(rng.uniform(-0.7, 0.7, size=K))
        ww, hh = s * np.sqrt(asp), s / np.sqrt(asp)
    cx = rng.uniform(0, img_w, size=K)
    cy = rng.uniform(0, img_h, size=K)
    # sub-pixel jitter so that sample points are generic (not on pixel centres)
    x1 = np.clip(cx - ww / 2, 0, img_w)
    x2 = np.clip(cx + ww / 2, 0, img_w)
    y1 = np.clip(cy - hh / 2, 0, img_h)
    y2 = np.clip(cy + hh / 2, 0, img_h)
    rois = np.stack([rng.randint(0, batch, size=K).astype(np.float64), x1, y1, x2, y2], axis=1).astype(np.float32)
    if edge_cases and K >= 8:
        rois[0, 1:] = [10.0, 10.0, 10.0, 10.0]                    # zero area
        rois[1, 1:] = [-40.0, -30.0, 20.0, 25.0]                  # partially outside (top-left)
        rois[2, 1:] = [img_w - 10.0, img_h - 12.0, img_w + 50.0, img_h + 40.0]  # partially outside
        rois[3, 1:] = [img_w + 20.0, img_h + 20.0, img_w + 90.0, img_h + 70.0]  # fully outside
        rois[4, 1:] = [0.0, 0.0, float(img_w), float(img_h)]      # whole image
        rois[5, 1:] = [100.25, 50.75, 103.0, 52.5]                # tiny (< 1 feature px)
        rois[6, 1:] = [30.0, 40.0, 25.0, 35.0]                    # inverted
        rois[7, 1:] = [-200.0, -200.0, -100.0, -100.0]            # fully outside (negative)
    return rois


def make_mask_probs(N, M, seed):
    """Mask-head probabilities [N,M,M]: sigmoid(N(0,2)) smoothed 3x3 (few pixels sit near 0.5)."""
    rng = np.random.RandomState(seed)
    z = rng.normal(0.0, 2.0, size=(N, M + 2, M + 2))
    sm = np.zeros((N, M, M))
    for dy in range(3):
        for dx in range(3):
            sm += z[:, dy:dy + M, dx:dx + M]
    sm /= 3.0
    return _sigmoid64(sm).astype(np.float32)


def make_box_scores(shape, seed):
    """Box-head class-1 probabilities U(0,1) (stand-in for softmax(cls_logits)[:,1],
    src/custom_maskrcnn.py:182-183)."""
    rng = np.random.RandomState(seed)
    return rng.uniform(0.0, 1.0, size=shape).astype(np.float32)


def make_det_boxes(N, seed, img_h=IMG_H, img_w=IMG_W, lo=16.0, hi=76.0, edge_cases=False):
    """Detection boxes [N,4] for paste tests: 16-76 px cells, fractional coords."""
    rng = np.random.RandomState(seed)
    ww = rng.uniform(lo, hi, size=N)
    hh = rng.uniform(lo, hi, size=N)
    x1 = rng.uniform(-8.0, img_w - 8.0, size=N)
    y1 = rng.uniform(-8.0, img_h - 8.0, size=N)
    b = np.stack([x1, y1, x1 + ww, y1 + hh], axis=1).astype(np.float32)
    if edge_cases and N >= 6:
        b[0] = [5.9, 7.2, 5.95, 30.0]                             # zero width after truncation
        b[1] = [-20.0, -20.0, -3.0, -1.0]                         # fully outside
        b[2] = [0.0, 0.0, float(img_w), float(img_h)]             # whole frame
        b[3] = [img_w - 9.5, img_h - 7.5, img_w + 30.0, img_h + 30.0]  # clipped at bottom-right
        b[4] = [12.0, 12.0, 14.0, 13.0]                           # 2x1 px
        b[5] = [-0.9, -0.9, 10.2, 9.7]                            # negative coords truncate toward 0
    return b

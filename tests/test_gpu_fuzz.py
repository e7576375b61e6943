"""GPU parity, randomized: every kernel family of the path against the CPU oracle on seeded random shapes and
parameters (odd sizes, ragged counts, boxes hanging over the frame, tiny and huge RoIs), through the C-ABI.
Bit-exact for indices / keep lists / clipped boxes / masks; 1e-5 relative for RoIAlign; 1e-6 absolute for scores."""
import numpy as np
import pytest
import torch

from gpu_util import N, T, assert_close_rel, nhwc

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("roi_cpu_coords")]   # oracle / golden = torchvision's CPU-op coordinate rule


@pytest.fixture(scope="module")
def ops():
    from livecell_instance_segmentation_b200 import ops as o
    return o


def random_boxes(rng, n, H, W, lo=2.0, hi=None, overhang=True):
    hi = hi or max(H, W) * 0.6
    w = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    h = np.exp(rng.uniform(np.log(lo), np.log(hi), n))
    cx = rng.uniform(-10 if overhang else 0, W + (10 if overhang else 0), n)
    cy = rng.uniform(-10 if overhang else 0, H + (10 if overhang else 0), n)
    b = np.stack([cx - w / 2, cy - h / 2, cx + w / 2, cy + h / 2], 1).astype(np.float32)
    if not overhang:
        b[:, [0, 2]] = b[:, [0, 2]].clip(0, W)
        b[:, [1, 3]] = b[:, [1, 3]].clip(0, H)
    return b


@pytest.mark.parametrize("seed", range(12))
def test_nms_random(ops, oracle, seed):
    rng = np.random.RandomState(1000 + seed)
    S = int(rng.randint(1, 5))
    stride = int(rng.choice([7, 40, 255, 256, 257, 600, 1023, 1500, 2048, 2100, 3000]))
    H, W = 520, 704
    dense = rng.rand() < 0.5     # dense: many overlaps (small frame region), long suppression chains
    boxes = np.stack([random_boxes(rng, stride, 120 if dense else H, 160 if dense else W, lo=8, hi=90) for _ in range(S)])
    scores = rng.rand(S, stride).astype(np.float32)
    if rng.rand() < 0.3:         # blocks of tied scores
        scores = np.round(scores * 8) / 8
    counts = rng.randint(0, stride + 1, size=S).astype(np.int32)
    counts[rng.randint(S)] = stride
    thr = float(rng.choice([0.1, 0.4, 0.5, 0.7, 0.95]))
    post_n = int(rng.choice([1, 10, stride // 2 + 1, stride]))
    use_thr = bool(rng.rand() < 0.5)
    keep, kc = ops.nms_batched(T(boxes), T(scores), thr, post_n=post_n, counts=T(counts), score_thresh=0.4 if use_thr else None)
    keep, kc = N(keep), N(kc)
    for s in range(S):
        c = counts[s]
        ref = oracle.nms(boxes[s, :c], scores[s, :c], thr, score_thresh=0.4, use_score_thresh=use_thr, post_n=post_n)
        assert kc[s] == len(ref), (seed, s)
        assert np.array_equal(keep[s, : kc[s]], ref), (seed, s)


@pytest.mark.parametrize("seed", range(10))
def test_roi_align_random(ops, oracle, seed):
    rng = np.random.RandomState(2000 + seed)
    P = int(rng.choice([7, 14]))
    C = int(rng.choice([4, 36, 64, 100, 256]))
    Nb = int(rng.randint(1, 4))
    h, w = int(rng.randint(5, 70)), int(rng.randint(5, 90))
    scale = float(rng.choice([0.25, 0.125, 0.5]))
    H, W = h / scale, w / scale
    K = int(rng.choice([1, 3, 33, 130]))
    boxes = random_boxes(rng, K, H, W, lo=1.0, hi=max(H, W) * 1.2)
    boxes[0] = [W * 0.3, H * 0.3, W * 0.3, H * 0.3]                     # degenerate (w = h = 0)
    if K > 2:
        boxes[1] = [-3 * W, -3 * H, -2 * W, -2 * H]                     # completely outside
        boxes[2] = [-5, -5, W + 5, H + 5]                               # larger than the map
    bidx = rng.randint(0, Nb, size=K).astype(np.float32)
    if K > 3:
        bidx[3] = -1                                                     # padding row
    rois = np.concatenate([bidx[:, None], boxes], 1).astype(np.float32)
    feat = rng.randn(Nb, C, h, w).astype(np.float32)
    layout = rng.choice(["nchw", "nhwc"])
    f = T(feat) if layout == "nchw" else nhwc(T(feat))
    out = ops.roi_align_fwd([f], [scale], T(rois), None, (P, P), 2, False)
    ref = oracle.roi_align_fwd(feat, rois, P, P, scale, 2, False)
    assert_close_rel(N(out), ref, 1e-5)
    gout = rng.randn(K, C, P, P).astype(np.float32)
    gin = torch.empty_like(f)
    ops.roi_align_bwd(T(gout), [gin], [scale], T(rois), None, 2, False, zero_grad=True)
    assert_close_rel(N(gin), oracle.roi_align_bwd(gout, rois, feat.shape, scale, 2, False), 1e-5)


@pytest.mark.parametrize("seed", range(10))
def test_paste_random(ops, oracle, synth, seed):
    rng = np.random.RandomState(3000 + seed)
    H = int(rng.choice([33, 64, 222, 256, 520]))
    W = int(rng.choice([48, 80, 300, 256, 704, 100]))       # 300 and 100: not multiples of 16 -> generic kernel
    M = int(rng.choice([14, 28]))
    n = int(rng.randint(1, 40))
    boxes = random_boxes(rng, n, H, W, lo=1.0, hi=max(H, W) * 1.5)
    boxes[0] = [5.9, 5.9, 5.95, 30]                          # empty after truncation
    if n > 1:
        boxes[1] = [-20, -20, W + 20, H + 20]               # the whole frame
    probs = synth.make_mask_probs(n, M, 3100 + seed)
    valid = (rng.rand(n) < 0.8).astype(np.uint8)
    thr = float(rng.choice([0.5, 0.3]))
    out = torch.full((n, H, W), 7, dtype=torch.uint8, device="cuda:0")
    ops.paste_masks(T(probs), T(boxes), H, W, thr, 255, valid=T(valid), out=out)
    ref = np.full((n, H, W), 7, np.uint8)
    oracle.paste_masks(probs, boxes, H, W, thr, 255, valid=valid, out=ref)
    assert np.array_equal(N(out), ref), seed


@pytest.mark.parametrize("seed", range(10))
def test_rpn_select_random(ops, oracle, synth, seed):
    rng = np.random.RandomState(4000 + seed)
    B = int(rng.randint(1, 4))
    h, w = int(rng.randint(3, 60)), int(rng.randint(3, 80))
    A = 9
    k = int(rng.choice([1, 17, 250, 500, 2000]))
    img_h, img_w = 4 * h + int(rng.randint(0, 4)), 4 * w + int(rng.randint(0, 4))
    n_cells = int(rng.randint(0, max(1, h * w // 3)))
    obj = synth.make_objectness(B, A, h, w, n_cells=n_cells, seed=4100 + seed, k=min(k, A * h * w))
    thr = float(rng.choice([0.01, 0.3, 0.9]))
    min_size = float(rng.choice([0.0, 5.0, 10.0, 40.0]))
    base = oracle.base_anchors()
    boxes, scores, index, counts = ops.rpn_select([T(obj)], k=k, img_size=(img_h, img_w), score_thresh=thr, min_size=min_size,
                                                  strides=[4], base=base)
    boxes, scores, index, counts = N(boxes), N(scores), N(index), N(counts)
    for b in range(B):
        rb, rs, ri = oracle.rpn_select(obj[b], base=base, stride=4, k=k, score_thresh=thr, min_size=min_size, img_h=img_h, img_w=img_w)
        c = counts[b, 0]
        assert c == len(ri), (seed, b)
        assert np.array_equal(index[b, 0, :c], ri), (seed, b)
        assert np.array_equal(boxes[b, 0, :c], rb), (seed, b)
        assert np.abs(scores[b, 0, :c] - rs).max(initial=0.0) <= 1e-6, (seed, b)


@pytest.mark.parametrize("seed", range(6))
def test_box_iou_and_mask_targets_random(ops, oracle, seed):
    rng = np.random.RandomState(5000 + seed)
    n, g = int(rng.choice([1, 50, 3000])), int(rng.choice([1, 7, 40]))
    H, W = 256, 300
    a, b = random_boxes(rng, n, H, W), random_boxes(rng, g, H, W, overhang=False)
    riou, rv, ri = oracle.box_iou(a, b)
    assert np.array_equal(N(ops.box_iou(T(a), T(b))), riou)
    vals, idx = ops.box_iou_max(T(a), T(b))
    assert np.array_equal(N(vals), rv) and np.array_equal(N(idx), ri)
    masks = (rng.rand(g, H, W) < 0.5).astype(np.uint8)
    gi = rng.randint(0, g, size=n).astype(np.int64)
    inside = random_boxes(rng, n, H, W, lo=4.0, hi=120.0, overhang=False)
    got = ops.mask_targets(T(masks), T(inside), T(gi), 28)
    assert np.abs(N(got) - oracle.mask_targets(masks, inside, gi, 28)).max() <= 1e-6

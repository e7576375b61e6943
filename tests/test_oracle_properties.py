"""Size-independent properties of the CPU oracle (oracle/lcr_oracle.c) — the checker the GPU parity tests lean on is itself
checked here against DEFINITIONS that share no code with it (numpy restatements of what each reference call means), beyond
the reference-generated fixtures of tests/test_oracle_golden.py:

* greedy NMS (torchvision.ops.nms at src/utils/proposal_utils.py:55, src/custom_maskrcnn.py:192): the keep list is the
  sequential greedy definition on the fp32 IoU matrix, idempotent, and independent of the input order of untied boxes;
* RoIAlign backward is the adjoint of forward (<fwd(x), g> == <x, bwd(g)>), forward is linear, a constant map pools to
  the constant inside the map, and sum(bwd(1)) == K*C*PH*PW for RoIs inside the map (the bench's conservation check);
* paste (src/utils/mask_utils.py:129-171): pixels outside the integer box stay zero, an all-ones mask fills the clipped
  box exactly, the batch equals per-detection calls;
* the fused matcher agrees with thresholds applied to the full IoU matrix.
"""
import numpy as np
import pytest


def _iou_matrix(b):
    """torchvision.ops.box_iou in fp32, operation for operation (TV:ops/boxes.py:308-370)."""
    b = b.astype(np.float32)
    area = (b[:, 2] - b[:, 0]) * (b[:, 3] - b[:, 1])
    lt = np.maximum(b[:, None, :2], b[None, :, :2])
    rb = np.minimum(b[:, None, 2:], b[None, :, 2:])
    wh = np.clip(rb - lt, 0, None).astype(np.float32)
    inter = wh[..., 0] * wh[..., 1]
    with np.errstate(invalid="ignore", divide="ignore"):
        return inter / ((area[:, None] + area[None, :]) - inter)


def _greedy(b, s, thr):
    order = np.argsort(-s, kind="stable")
    iou = _iou_matrix(b)
    alive = np.ones(len(b), bool)
    keep = []
    for i in order:
        if not alive[i]:
            continue
        keep.append(i)
        alive &= ~(iou[i] > thr)          # torchvision CPU rule: fp32 IoU against the double threshold
        alive[i] = False
    return np.asarray(keep, np.int64)


@pytest.mark.parametrize("seed,n,thr", [(0, 300, 0.5), (1, 1200, 0.4), (2, 64, 0.7), (3, 1, 0.5)])
def test_nms_is_the_greedy_definition_and_idempotent(oracle, synth, seed, n, thr):
    rng = np.random.RandomState(seed)
    b = synth.make_det_boxes(n, 40 + seed, lo=24.0, hi=120.0)
    s = rng.permutation(n).astype(np.float32) / n                 # untied scores
    keep = oracle.nms(b, s, thr)
    assert np.array_equal(keep, _greedy(b, s, thr))
    iou = _iou_matrix(b[keep])
    np.fill_diagonal(iou, 0)
    assert not (iou > thr).any()                                  # survivors do not suppress each other
    assert np.array_equal(oracle.nms(b[keep], s[keep], thr), np.arange(len(keep)))   # idempotent, order kept
    perm = rng.permutation(n)                                     # input order is irrelevant without ties
    assert np.array_equal(perm[oracle.nms(b[perm], s[perm], thr)], keep)
    assert np.array_equal(oracle.nms(b, s, thr, post_n=min(5, n)), keep[:5])         # post_n only truncates


@pytest.mark.parametrize("P,sr,aligned", [(7, 2, False), (14, 2, True), (5, 0, False)])
def test_roi_align_backward_is_the_adjoint_of_forward(oracle, synth, P, sr, aligned):
    rng = np.random.RandomState(7)
    N, C, H, W = 2, 6, 33, 41
    x = rng.randn(N, C, H, W).astype(np.float32)
    rois = synth.make_rois(40, 3, img_h=H * 4, img_w=W * 4, batch=N, edge_cases=True)
    y = oracle.roi_align_fwd(x, rois, P, P, 0.25, sr, aligned)
    g = rng.randn(*y.shape).astype(np.float32)
    gx = oracle.roi_align_bwd(g, rois, x.shape, 0.25, sr, aligned)
    lhs, rhs = float((y.astype(np.float64) * g).sum()), float((x.astype(np.float64) * gx).sum())
    assert abs(lhs - rhs) <= 1e-4 * max(abs(lhs), abs(rhs), 1.0), (lhs, rhs)
    # linearity of forward
    x2 = rng.randn(N, C, H, W).astype(np.float32)
    y2 = oracle.roi_align_fwd(x2, rois, P, P, 0.25, sr, aligned)
    y12 = oracle.roi_align_fwd((2 * x + x2).astype(np.float32), rois, P, P, 0.25, sr, aligned)
    assert np.abs(y12 - (2 * y + y2)).max() <= 1e-4


def test_roi_align_constant_map_and_gradient_mass(oracle):
    N, C, H, W, P = 1, 3, 40, 52, 7
    rng = np.random.RandomState(11)
    x1, y1 = rng.uniform(8, 90, 30), rng.uniform(8, 70, 30)
    rois = np.stack([np.zeros(30), x1, y1, x1 + rng.uniform(12, 100, 30), y1 + rng.uniform(12, 80, 30)], 1).astype(np.float32)
    assert rois[:, 3].max() * 0.25 < W - 1 and rois[:, 4].max() * 0.25 < H - 1           # every sample inside the map
    y = oracle.roi_align_fwd(np.full((N, C, H, W), 3.25, np.float32), rois, P, P, 0.25, 2, False)
    assert np.abs(y - 3.25).max() <= 1e-5                                                # bilinear weights sum to 1
    gx = oracle.roi_align_bwd(np.ones((30, C, P, P), np.float32), rois, (N, C, H, W), 0.25, 2, False)
    assert abs(float(gx.astype(np.float64).sum()) - 30 * C * P * P) <= 1e-3 * 30 * C * P * P   # bench's sum_check
    assert (gx >= 0).all()


def test_paste_is_confined_to_the_integer_box_and_fills_it(oracle, synth):
    H, W, M = 96, 128, 28
    boxes = synth.make_det_boxes(24, 9, img_h=H, img_w=W, edge_cases=True)
    ones = np.ones((24, M, M), np.float32)
    out = oracle.paste_masks(ones, boxes, H, W)
    assert out.dtype == np.uint8 and set(np.unique(out)) <= {0, 255}
    for i, (x1, y1, x2, y2) in enumerate(boxes):
        # mask_utils.py:147-160: int() truncation toward zero, then clamped to the frame; empty boxes paste nothing
        ix1, iy1, ix2, iy2 = max(int(x1), 0), max(int(y1), 0), min(int(x2), W), min(int(y2), H)
        inside = np.zeros((H, W), bool)
        if ix2 > ix1 and iy2 > iy1:
            inside[iy1:iy2, ix1:ix2] = True
        assert not out[i][~inside].any(), i                       # nothing outside the box
        assert (out[i][inside] == 255).all(), i                   # probability 1 everywhere > 0.5: the box is filled
    probs = synth.make_mask_probs(24, M, 5)
    batch = oracle.paste_masks(probs, boxes, H, W)
    for i in range(24):
        assert np.array_equal(batch[i], oracle.paste_masks(probs[i:i + 1], boxes[i:i + 1], H, W)[0])
    assert not oracle.paste_masks(np.zeros((24, M, M), np.float32), boxes, H, W).any()
    valid = np.arange(24) % 2 == 0                                # invalid (padding) detections leave their frame zero
    half = oracle.paste_masks(probs, boxes, H, W, valid=valid)
    assert np.array_equal(half[valid], batch[valid]) and not half[~valid].any()


def test_matcher_equals_thresholds_on_the_full_matrix(oracle, synth):
    a = synth.make_det_boxes(3000, 21, lo=16.0, hi=140.0, edge_cases=True)
    g = synth.make_det_boxes(70, 22, lo=16.0, hi=90.0)
    iou, mx, am = oracle.box_iou(a, g)
    ref = _iou_matrix(np.concatenate([a, g]))[: len(a), len(a):]
    assert np.array_equal(iou, ref, equal_nan=True)               # the C restatement against the numpy definition
    assert np.array_equal(mx, ref.max(1)) and np.array_equal(am, ref.argmax(1))
    for pos_thr, neg_thr in ((0.5, 0.3), (0.4, None)):
        _, _, pos, neg, cnt = oracle.match_boxes(a, g, pos_thr, neg_thr)
        lo = pos_thr if neg_thr is None else neg_thr
        assert np.array_equal(pos, ref.max(1) >= np.float32(pos_thr)) and np.array_equal(neg, ref.max(1) < np.float32(lo))
        assert cnt.tolist() == [int(pos.sum()), int(neg.sum())] and not (pos & neg).any()
        if neg_thr is None:
            assert (pos | neg).all()                              # complement thresholds partition the rows (no NaN here)

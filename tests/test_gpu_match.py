"""GPU parity for the training-side siblings: box IoU / fused row max and batched
mask targets, through the C-ABI, against the reference-generated golden vectors and the oracle."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    from livecell_instance_segmentation_b200 import ops as o
    return o


def test_box_iou_golden(ops, golden):
    from gpu_util import N, T
    g = golden("match")
    iou = N(ops.box_iou(T(g["boxes"]), T(g["gt"])))
    assert np.array_equal(iou, g["iou"], equal_nan=True)                 # bit-exact incl. NaN (0/0) and -0/+0 rows
    mx, am = ops.box_iou_max(T(g["boxes"]), T(g["gt"]))
    assert np.array_equal(N(mx), g["max_iou"], equal_nan=True)
    assert np.array_equal(N(am), g["argmax"]) and am.dtype == torch.int64


def test_box_iou_full_size_vs_oracle_and_torchvision(ops, oracle, synth):
    """C1 geometry: 205 920 anchors x 160 ground-truth boxes (gt count spans two shared-memory chunks at 1500)."""
    import torchvision
    from gpu_util import N, T
    anc = N(ops.anchors(130, 176, 4, ops.base_anchors(), "cuda:0"))
    for G, seed in ((160, 3), (1500, 4)):
        gt = synth.make_det_boxes(G, seed)
        _, mx_ref, am_ref = oracle.box_iou(anc, gt, want_matrix=False)
        mx, am = ops.box_iou_max(T(anc), T(gt))
        assert np.array_equal(N(mx), mx_ref, equal_nan=True) and np.array_equal(N(am), am_ref)
        tv_mx, tv_am = torchvision.ops.box_iou(T(anc), T(gt)).max(dim=1)
        assert np.array_equal(N(mx), N(tv_mx), equal_nan=True)
        # on exact ties the CUDA reduction of torch.max may pick another index: compare through the values
        assert np.array_equal(N(torchvision.ops.box_iou(T(anc), T(gt)).gather(1, am[:, None])[:, 0]), N(tv_mx), equal_nan=True)
    # thresholds used by the reference on top of the row max (rpn.py:75-76, custom_maskrcnn.py:225)
    pos, neg = (mx >= 0.5), (mx < 0.3)
    assert int(pos.sum()) == int((torch.from_numpy(mx_ref) >= 0.5).sum()) and int(neg.sum()) == int((torch.from_numpy(mx_ref) < 0.3).sum())


def test_match_boxes_golden_oracle_and_torchvision(ops, oracle, golden, synth):
    """lcr_match_boxes_f32 — row max + threshold masks + sums in one kernel (src/components/rpn.py:72-81: 0.5 / 0.3;
    src/custom_maskrcnn.py:221-225, 249-251: 0.4).  Bit-exact against (i) torch's own expressions on the reference-generated
    row max of the golden fixture (NaN rows, a tie), (ii) the oracle at C1 size across two ground-truth chunks, (iii) the
    reference's op chain run by torchvision / ATen on this GPU."""
    import torchvision
    from gpu_util import N, T
    g = golden("match")
    tmx = torch.from_numpy(g["max_iou"])
    for pos_thr, neg_thr in ((0.5, 0.3), (0.4, None)):
        mx, am, pos, neg, cnt = ops.match_boxes(T(g["boxes"]), T(g["gt"]), pos_thr, neg_thr)
        assert pos.dtype == torch.bool and neg.dtype == torch.bool and cnt.dtype == torch.int32
        assert np.array_equal(N(mx), g["max_iou"], equal_nan=True) and np.array_equal(N(am), g["argmax"])
        t_pos, t_neg = tmx >= pos_thr, tmx < (pos_thr if neg_thr is None else neg_thr)
        assert np.array_equal(N(pos), t_pos.numpy()) and np.array_equal(N(neg), t_neg.numpy())
        assert cnt.tolist() == [int(t_pos.sum()), int(t_neg.sum())]
    anc = ops.anchors(130, 176, 4, ops.base_anchors(), "cuda:0")
    for G, seed in ((160, 3), (1500, 4)):
        gt = synth.make_det_boxes(G, seed)
        r_mx, r_am, r_pos, r_neg, r_cnt = oracle.match_boxes(N(anc), gt, 0.5, 0.3)
        mx, am, pos, neg, cnt = ops.match_boxes(anc, T(gt), 0.5, 0.3)
        assert np.array_equal(N(mx), r_mx, equal_nan=True) and np.array_equal(N(am), r_am)
        assert np.array_equal(N(pos), r_pos) and np.array_equal(N(neg), r_neg) and cnt.tolist() == r_cnt.tolist()
        assert 0 < r_cnt[0] < r_cnt[1] < anc.shape[0]                       # both masks populated, some anchors ignored
        tv_mx, _ = torchvision.ops.box_iou(anc, T(gt)).max(dim=1)           # the chain the reference runs (rpn.py:72-77)
        assert torch.equal(pos, tv_mx >= 0.5) and torch.equal(neg, tv_mx < 0.3)
        assert cnt.tolist() == [int((tv_mx >= 0.5).sum().item()), int((tv_mx < 0.3).sum().item())]
        # a second call on the same stream zeroes the counters itself
        assert ops.match_boxes(anc, T(gt), 0.5, 0.3)[4].tolist() == r_cnt.tolist()
    # a value exactly on a threshold: IoU of [0,0,2,2] with [0,0,2,1] is exactly 0.5 -> positive; 0.25 is < 0.3
    b = T(np.array([[0, 0, 2, 1], [0, 0, 1, 1], [5, 5, 6, 6], [3, 3, 3, 3]], np.float32))
    mx, am, pos, neg, cnt = ops.match_boxes(b, T(np.array([[0, 0, 2, 2], [3, 3, 3, 3]], np.float32)), 0.5, 0.3)
    assert N(mx)[:3].tolist() == [0.5, 0.25, 0.0] and np.isnan(N(mx)[3])
    assert N(pos).tolist() == [True, False, False, False] and N(neg).tolist() == [False, True, True, False]
    assert cnt.tolist() == [1, 2]


@pytest.mark.parametrize("seed", range(4))
def test_match_boxes_short_path_is_exact_on_hostile_boxes(ops, oracle, seed):
    """Written for a short-path variant of match_kernel (no iou_tv for pairs with an empty intersection and a finite positive
    area sum: bit-identical here, measured slower, reverted — profiles/r02d_match_ncu.md) and kept for the kernel in use:
    the result must be the bits of lcr_box_iou_max_f32 (pinned to the golden row max) and of the oracle for malformed
    input too: inverted boxes (negative
    areas -> -0 IoUs), zero-area boxes (0/0 = NaN), 1e38-sized and infinite coordinates, NaN coordinates, dense
    overlaps, and a ground-truth list spanning two shared-memory chunks."""
    from gpu_util import N, T
    rng = np.random.RandomState(900 + seed)
    n, g = 4000, (37, 1100, 300, 5)[seed]

    def boxes(k, nonfinite):
        x1, y1 = rng.uniform(-50, 400, k), rng.uniform(-50, 400, k)
        b = np.stack([x1, y1, x1 + rng.uniform(0, 200, k), y1 + rng.uniform(0, 200, k)], 1).astype(np.float32)
        kind = rng.randint(0, 12, k)
        b[kind == 0, 2] = b[kind == 0, 0] - rng.uniform(1, 30, (kind == 0).sum())        # inverted in x: negative area
        b[kind == 1, 3] = b[kind == 1, 1]                                                # zero height
        b[kind == 2] = b[kind == 2][:, [0, 1, 0, 1]]                                     # a point: zero area
        b[kind == 3, 2:] = b[kind == 3, :2] - 5.0                                        # inverted in both: positive area, empty
        if nonfinite:
            b[kind == 4] *= np.float32(1e36)                                             # areas overflow to inf
            b[kind == 5, 2] = np.inf
            b[kind == 6, 0] = -np.inf
            b[kind == 7, rng.randint(0, 4)] = np.nan
        return b
    a, gt = boxes(n, seed >= 1), boxes(g, seed == 3)                # non-finite ground truth poisons every row: one small case only
    with np.errstate(all="ignore"):
        r_mx, r_am, r_pos, r_neg, r_cnt = oracle.match_boxes(a, gt, 0.5, 0.3)
    k_mx, k_am = ops.box_iou_max(T(a), T(gt))
    mx, am, pos, neg, cnt = ops.match_boxes(T(a), T(gt), 0.5, 0.3)
    assert np.array_equal(N(mx).view(np.uint32), N(k_mx).view(np.uint32))       # bit for bit, -0 and NaN rows included
    assert torch.equal(am, k_am)
    assert np.array_equal(N(mx), r_mx, equal_nan=True) and np.array_equal(N(am), r_am)
    assert np.array_equal(N(pos), r_pos) and np.array_equal(N(neg), r_neg) and cnt.tolist() == r_cnt.tolist()
    assert np.isnan(N(mx)).any() and (seed == 3 or (np.isfinite(N(mx)) & (N(mx) > 0)).sum() > 1000)


def test_match_boxes_edge_cases(ops):
    from gpu_util import T
    from livecell_instance_segmentation_b200._lib import LcrError
    g = T(np.array([[0, 0, 1, 1]], np.float32))
    mx, am, pos, neg, cnt = ops.match_boxes(T(np.zeros((0, 4), np.float32)), g, 0.5, 0.3)
    assert mx.shape == (0,) and pos.shape == (0,) and cnt.tolist() == [0, 0]
    with pytest.raises(LcrError):
        ops.match_boxes(g, T(np.zeros((0, 4), np.float32)), 0.5, 0.3)   # no ground truth: the reference never matches
    with pytest.raises(LcrError):
        ops.match_boxes(g.cpu(), g.cpu(), 0.5)                          # CPU tensors: no fallback


def test_box_iou_edge_cases(ops):
    from gpu_util import T
    from livecell_instance_segmentation_b200._lib import LcrError
    a = T(np.zeros((0, 4), np.float32))
    g = T(np.array([[0, 0, 1, 1]], np.float32))
    assert tuple(ops.box_iou(a, g).shape) == (0, 1)
    mx, am = ops.box_iou_max(a, g)
    assert mx.shape == (0,) and am.shape == (0,)
    with pytest.raises(LcrError):
        ops.box_iou_max(g, a)                       # no ground truth: torch.max over an empty dim raises too
    with pytest.raises(LcrError):
        ops.box_iou(g.cpu(), g.cpu())               # CPU tensors: no fallback


def test_mask_targets_golden_and_oracle(ops, oracle, golden, synth):
    from gpu_util import N, T
    g = golden("match")
    tg = N(ops.mask_targets(T(g["masks"]), T(g["t_boxes"]), T(g["t_index"]), 28))
    np.testing.assert_allclose(tg, g["targets"], rtol=0, atol=1e-6)       # reference (ATen CPU bilinear)
    assert np.array_equal(tg, oracle.mask_targets(g["masks"], g["t_boxes"], g["t_index"], 28))   # same expression tree
    # full 704x520 frames, 300 positives matched to 40 cells, M = 28 and 14
    rng = np.random.RandomState(5)
    H, W, G, K = 520, 704, 40, 300
    gb = synth.make_det_boxes(G, 6)
    yy, xx = np.mgrid[0:H, 0:W]
    masks = np.stack([((xx >= b[0]) & (xx < b[2]) & (yy >= b[1]) & (yy < b[3])).astype(np.uint8) for b in gb])
    idx = rng.randint(0, G, size=K).astype(np.int64)
    boxes = (gb[idx] + rng.uniform(-4, 4, size=(K, 4))).astype(np.float32)
    for M in (28, 14):
        out = N(ops.mask_targets(T(masks), T(boxes), T(idx), M))
        assert np.array_equal(out, oracle.mask_targets(masks, boxes, idx, M))
    # default index (target k <- mask k) and an out-of-range index (all zeros)
    out = N(ops.mask_targets(T(masks), T(gb), None, 28))
    assert np.array_equal(out, oracle.mask_targets(masks, gb, None, 28))
    bad = idx.copy()
    bad[7] = -1
    out = N(ops.mask_targets(T(masks), T(boxes), T(bad), 28))
    assert float(np.abs(out[7]).max()) == 0.0


def test_mask_loss_shim_matches_reference_formula(ops, synth):
    """compute_mask_loss_from_gt (mask_utils.py:49-126) through the shim == the reference's formula evaluated with
    torch/torchvision ops on the same device (IoU match > 0.3, per-positive bilinear targets, BCE on class 1)."""
    import torch.nn.functional as F
    import torchvision
    from gpu_util import T
    from livecell_instance_segmentation_b200.src.utils.mask_utils import compute_mask_loss_from_gt, extract_mask_target
    rng = np.random.RandomState(17)
    H, W, G, K = 256, 256, 12, 96
    gb = synth.make_det_boxes(G, 18, img_h=H, img_w=W, lo=20, hi=50)
    yy, xx = np.mgrid[0:H, 0:W]
    masks = np.stack([((xx >= b[0]) & (xx < b[2]) & (yy >= b[1]) & (yy < b[3])).astype(np.uint8) for b in gb])
    props = np.concatenate([gb[rng.randint(0, G, size=K // 2)] + rng.uniform(-5, 5, size=(K // 2, 4)),
                            synth.make_det_boxes(K // 2, 19, img_h=H, img_w=W, lo=20, hi=50)]).astype(np.float32)
    logits = T(rng.standard_normal((K, 2, 28, 28)).astype(np.float32))
    targets = [{"boxes": T(gb[:7]), "masks": T(masks[:7]), "labels": torch.ones(7)},
               {"boxes": T(gb[7:]), "masks": T(masks[7:]), "labels": torch.ones(G - 7)}]
    loss = compute_mask_loss_from_gt(logits, T(props), targets, "cuda:0")
    # the reference formula, with ATen/torchvision CUDA ops
    ious = torchvision.ops.box_iou(T(props), T(gb))
    mx, idx = ious.max(dim=1)
    pos = mx > 0.3
    assert int(pos.sum()) > 10
    tg = []
    for gi in idx[pos].tolist():
        x1, y1, x2, y2 = [int(v) for v in gb[gi]]
        x1 = max(0, min(x1, W - 1)); y1 = max(0, min(y1, H - 1)); x2 = max(x1 + 1, min(x2, W)); y2 = max(y1 + 1, min(y2, H))
        crop = T(masks[gi])[y1:y2, x1:x2].float()[None, None]
        tg.append(F.interpolate(crop, size=(28, 28), mode="bilinear", align_corners=False)[0, 0])
    ref = F.binary_cross_entropy_with_logits(logits[pos][:, 1], torch.stack(tg), reduction="mean")
    assert abs(float(loss) - float(ref)) <= 1e-6 * max(1.0, abs(float(ref)))
    t1 = extract_mask_target(T(masks[3]), T(gb[3]), 28)
    assert tuple(t1.shape) == (28, 28)


def test_mask_tail_golden_and_oracle(ops, oracle, golden):
    from gpu_util import N, T
    g = golden("tail_stitch")
    p = N(ops.mask_tail(T(g["logits14"]), 28, 1))
    np.testing.assert_allclose(p, g["probs28"], rtol=0, atol=1e-6)             # reference head tail + sigmoid
    np.testing.assert_allclose(p, oracle.mask_tail(g["logits14"], 28, 1), rtol=0, atol=2e-7)   # glibc vs CUDA expf
    np.testing.assert_allclose(N(ops.mask_tail(T(g["logits28"]), 28, 1)), g["probs28_same"], rtol=0, atol=1e-6)
    assert tuple(ops.mask_tail(T(np.zeros((0, 2, 14, 14), np.float32))).shape) == (0, 28, 28)
    # the fused tail feeds the paste kernel exactly like the unfused chain (thresholded masks identical)
    import torch.nn.functional as F
    rng = np.random.RandomState(4)
    logits = T((rng.standard_normal((40, 2, 14, 14)) * 3).astype(np.float32))
    from livecell_instance_segmentation_b200 import synth
    boxes = T(synth.make_det_boxes(40, 9))
    unfused = torch.sigmoid(F.interpolate(logits, size=(28, 28), mode="bilinear", align_corners=False)[:, 1])
    a = ops.paste_masks(ops.mask_tail(logits, 28, 1), boxes, 520, 704)
    b = ops.paste_masks(unfused, boxes, 520, 704)
    assert float((a != b).float().mean()) < 1e-6       # ulp-level probability differences may flip a pixel sitting on 0.5


def test_tile_stitch_filter_matches_reference(golden):
    """stitch.filter_detections_by_border_mini_tiles on GPU-resident predictions == the reference's own
    filter_detections_by_border_mini_tiles (visualize.py:174-257) executed on CPU (fixture: 25 tiles x 14 detections)."""
    from gpu_util import T
    from test_oracle_golden import _stitch_results
    from livecell_instance_segmentation_b200 import stitch
    g = golden("tail_stitch")
    results = [{"tile_num": r["tile_num"], "prediction": {"boxes": T(r["boxes"]), "scores": T(r["scores"]), "masks": T(r["masks"])}}
               for r in _stitch_results(g)]
    kept = stitch.filter_detections_by_border_mini_tiles(list(reversed(results)), score_threshold=0.5, mask_threshold=0.4)
    assert [d["tile_num"] for d in kept] == g["kept_tile"].tolist()
    assert np.array_equal(np.array([d["area_fraction"] for d in kept]), g["kept_fraction"])     # float64 bit-exact
    assert np.array_equal(np.array([d["score"] for d in kept]), g["kept_score"])
    np.testing.assert_allclose(np.array([d["box"] for d in kept]), g["kept_box"], rtol=0, atol=1e-4)
    assert [int(d["mask"].sum()) for d in kept] == g["kept_mask_sum"].tolist()
    assert kept[0]["mask"].is_cuda and kept[0]["mask"].dtype == torch.bool
    # float probability masks [N,1,h,w] (the transfer model's output format) take the same path
    res2 = [{"tile_num": r["tile_num"], "prediction": {"boxes": r["prediction"]["boxes"], "scores": r["prediction"]["scores"],
                                                       "masks": (r["prediction"]["masks"].float() / 255.0)[:, None]}} for r in results]
    kept2 = stitch.filter_detections_by_border_mini_tiles(res2)
    assert [d["tile_num"] for d in kept2] == g["kept_tile"].tolist()


def test_mask_region_counts_with_boxes(ops, oracle, synth):
    """Scanning only the box area gives the same counts as the whole frame (a pasted mask is zero outside its box)."""
    from gpu_util import N, T
    n, H, W = 50, 520, 704
    boxes = synth.make_det_boxes(n, 12, edge_cases=True)
    masks = ops.paste_masks(T(synth.make_mask_probs(n, 28, 13)), T(boxes), H, W)
    rng = np.random.RandomState(3)
    R = 5
    x0, y0 = rng.randint(0, W - 50, size=(n, R)), rng.randint(0, H - 50, size=(n, R))
    rects = np.stack([x0, y0, x0 + rng.randint(1, 300, size=(n, R)), y0 + rng.randint(1, 300, size=(n, R))], axis=-1)
    rects[..., 2] = np.minimum(rects[..., 2], W)
    rects[..., 3] = np.minimum(rects[..., 3], H)
    ro = np.arange(0, (n + 1) * R, R)
    t_ref, r_ref = oracle.mask_region_counts(N(masks), rects.reshape(-1, 4), ro)
    for bx in (None, T(boxes)):
        t, r = ops.mask_region_counts(masks, T(rects.reshape(-1, 4).astype(np.int32)), T(ro.astype(np.int32)), boxes=bx)
        assert np.array_equal(N(t), t_ref) and np.array_equal(N(r), r_ref)


def test_mask_region_counts_many_and_ragged_rect_lists(ops, oracle, synth):
    """more than 16 rectangles per detection used to be truncated silently; ragged lists (0, 1, 16, 17, 40
    rectangles) must all be counted."""
    from gpu_util import N, T
    n, H, W = 6, 222, 300
    boxes = synth.make_det_boxes(n, 21)
    boxes[:, [0, 2]] = boxes[:, [0, 2]].clip(0, W)
    boxes[:, [1, 3]] = boxes[:, [1, 3]].clip(0, H)
    boxes[4], boxes[5] = [40.0, 30.0, 120.0, 90.0], [100.5, 60.2, 180.0, 170.9]      # certainly inside the frame
    probs = synth.make_mask_probs(n, 28, 22)
    probs[4:] = 0.9                                                   # fully-on masks
    masks = ops.paste_masks(T(probs), T(boxes), H, W)
    rng = np.random.RandomState(5)
    per = [0, 1, 16, 17, 40, 33]
    ro = np.concatenate([[0], np.cumsum(per)])
    R = int(ro[-1])
    x0, y0 = rng.randint(0, W - 20, size=R), rng.randint(0, H - 20, size=R)
    rects = np.stack([x0, y0, np.minimum(x0 + rng.randint(1, 200, size=R), W), np.minimum(y0 + rng.randint(1, 200, size=R), H)], axis=-1)
    rects[ro[4] + 20] = rects[ro[5] + 32] = [0, 0, W, H]             # rectangles past the 16th that certainly see pixels
    t_ref, r_ref = oracle.mask_region_counts(N(masks), rects, ro)
    for bx in (None, T(boxes)):
        t, r = ops.mask_region_counts(masks, T(rects.astype(np.int32)), T(ro.astype(np.int32)), boxes=bx)
        assert np.array_equal(N(t), t_ref) and np.array_equal(N(r), r_ref)
    assert r_ref[ro[4] + 20] == t_ref[4] > 0 and r_ref[ro[5] + 32] == t_ref[5] > 0

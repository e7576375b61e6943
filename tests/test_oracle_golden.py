"""Pins the CPU oracle (oracle/lcr_oracle.c) to outputs of the reference itself.

The reference has no tests or golden vectors: tests/golden/*.npz were produced by
tests/golden/make_golden.py, which imports /root/reference and the installed torchvision CPU ops.
Bit-exact where the domain is integer/index/byte work; RoIAlign is bit-exact against the compiled
CPU op too (SURVEY App. B.1/B.2 restatement)."""
import hashlib

import numpy as np
import pytest


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def test_anchors_bit_exact(golden, oracle):
    g = golden("anchors")
    base = oracle.base_anchors(tuple(g["sizes"]), tuple(g["ratios"]))
    assert np.array_equal(oracle.anchors(5, 7, 4, base), g["small"])
    full = oracle.anchors(130, 176, 4, base)
    assert full.shape == (205920, 4)
    assert np.array_equal(sha(full), g["full_sha256"])
    assert np.array_equal(full[g["full_rows"]], g["full_vals"])
    assert np.array_equal(sha(oracle.anchors(56, 75, 4, base)), g["tile_sha256"])


def test_clip_filter_decode(golden, oracle):
    g = golden("box_utils")
    clipped = oracle.clip_boxes(g["boxes"], 520, 704)
    assert np.array_equal(clipped, g["clipped"])
    assert np.array_equal(oracle.filter_small_boxes(clipped, 10), g["keep10"])
    assert np.array_equal(oracle.filter_small_boxes(clipped, 5), g["keep5"])
    d1 = oracle.box_decode(g["deltas"], g["anchors"], (1, 1, 1, 1))
    d10 = oracle.box_decode(g["deltas"], g["anchors"], (10, 10, 5, 5))
    # expf differs between glibc and ATen's vectorised exp by ulps: 1e-5 relative (north star)
    np.testing.assert_allclose(d1, g["decoded_w1"], rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(d10, g["decoded_w10"], rtol=1e-5, atol=1e-4)
    # decode is the inverse of the reference's encode_boxes (src/utils/box_utils.py:4-28)
    rt = oracle.box_decode(g["encoded"], g["anchors"], (1, 1, 1, 1))
    np.testing.assert_allclose(rt, g["gt"], rtol=2e-5, atol=2e-4)


def _nms_after_select(oracle, obj, base, k, thr, nms_thr, post, min_size, H, W):
    boxes, scores, index = oracle.rpn_select(obj, base=base, stride=4, k=k, score_thresh=thr, min_size=min_size,
                                             img_h=H, img_w=W)
    keep = oracle.nms(boxes, None, nms_thr, post_n=post)
    return boxes, scores, index, keep


def test_proposals_small(golden, oracle):
    g = golden("proposals")
    base = oracle.base_anchors()
    for b in range(2):
        obj = g["small_obj"][b]
        boxes, scores, index, keep = _nms_after_select(oracle, obj, base, 100, 0.3, 0.4, 30, 10, 96, 128)
        # top-k index list (tie-free fixture): prefix that survives the score threshold
        ref_idx, ref_s = g[f"small_topk_index_{b}"], g[f"small_topk_scores_{b}"]
        b2, s2, i2 = oracle.rpn_select(obj, base=base, k=100, score_thresh=-1.0, min_size=0.0, img_h=96, img_w=128)
        assert np.array_equal(i2, ref_idx)
        np.testing.assert_allclose(s2, ref_s, rtol=0, atol=1e-6)
        assert np.array_equal(boxes[keep], g[f"small_inf_boxes_{b}"])
        np.testing.assert_allclose(scores[keep], g[f"small_inf_scores_{b}"], rtol=0, atol=1e-6)
        tb, _, _ = oracle.rpn_select(obj, base=base, k=120, score_thresh=0.01, min_size=5, img_h=96, img_w=128)
        assert np.array_equal(tb, g[f"small_train_boxes_{b}"])


@pytest.mark.parametrize("tag,seed,n_cells,k,post", [("c1", 21, 150, 250, 50), ("c3", 23, 2000, 2000, 1000)])
def test_proposals_full_size(golden, oracle, synth, tag, seed, n_cells, k, post):
    g = golden("proposals")
    assert int(g[f"{tag}_seed"]) == seed
    obj = synth.make_objectness(1, 9, 130, 176, n_cells=n_cells, seed=seed, k=k)[0]
    base = oracle.base_anchors()
    _, _, i2 = oracle.rpn_select(obj, base=base, k=k, score_thresh=-1.0, min_size=0.0)
    assert np.array_equal(i2, g[f"{tag}_topk_index"])
    boxes, scores, index, keep = _nms_after_select(oracle, obj, base, k, 0.3, 0.4, post, 10, 520, 704)
    assert np.array_equal(boxes[keep], g[f"{tag}_inf_boxes"])
    np.testing.assert_allclose(scores[keep], g[f"{tag}_inf_scores"], rtol=0, atol=1e-6)
    # gathered-anchor variant must agree with generated anchors
    anc = oracle.anchors(130, 176, 4, base)
    b3, s3, i3 = oracle.rpn_select(obj, anchors=anc, k=k, score_thresh=0.3, min_size=10)
    assert np.array_equal(b3, boxes) and np.array_equal(i3, index)


def test_training_proposals_c2(golden, oracle, synth):
    g = golden("proposals")
    obj = synth.make_objectness(1, 9, 64, 64, n_cells=20, seed=int(g["c2_seed"]), k=500)[0]
    base = oracle.base_anchors()
    tb, _, ti = oracle.rpn_select(obj, base=base, k=500, score_thresh=0.01, min_size=5, img_h=256, img_w=256)
    assert np.array_equal(tb, g["c2_train_boxes"])
    _, _, i2 = oracle.rpn_select(obj, base=base, k=500, score_thresh=-1.0, min_size=0.0, img_h=256, img_w=256)
    assert np.array_equal(i2, g["c2_topk_index"])


def test_nms_keep_lists(golden, oracle):
    g = golden("nms")
    for n in (250, 2000):
        for thr in (0.4, 0.5, 0.7):
            keep = oracle.nms(g[f"boxes_{n}"], g[f"scores_{n}"], thr)
            assert np.array_equal(keep, g[f"keep_{n}_{int(thr * 10)}"]), (n, thr)
    assert np.array_equal(oracle.nms(g["tie_boxes"], g["tie_scores"], 0.5), g["tie_keep"])
    assert np.array_equal(oracle.nms(g["eq_boxes"], g["eq_scores"], 0.4), g["eq_keep_04"])
    assert np.array_equal(oracle.nms(g["eq_boxes"], g["eq_scores"], 0.5), g["eq_keep_05"])
    assert np.array_equal(oracle.nms(g["zero_boxes"], g["zero_scores"], 0.4), g["zero_keep"])
    assert np.array_equal(oracle.nms(g["nan_boxes"], g["nan_scores"], 0.4), g["nan_keep"])


@pytest.mark.parametrize("tag,P,sr,al", [("p7", 7, 2, False), ("p14", 14, 2, False), ("p7a", 7, 2, True), ("p7ad", 7, 0, False)])
def test_roi_align_fwd_bwd(golden, oracle, tag, P, sr, al):
    g = golden("roi_align")
    out = oracle.roi_align_fwd(g["feat"], g["rois"], P, P, 0.25, sr, al)
    # restatement of SURVEY App. B.1: bit-identical to the compiled CPU op
    assert np.array_equal(out, g[f"out_{tag}"])
    gin = oracle.roi_align_bwd(g[f"gout_{tag}"], g["rois"], g["feat"].shape, 0.25, sr, al)
    np.testing.assert_allclose(gin, g[f"gin_{tag}"], rtol=1e-5, atol=1e-5)


def test_roi_align_list_form_and_multiscale(golden, oracle):
    g = golden("roi_align")
    rois = g["rois"].copy()
    rois[:, 0] = 0
    assert np.array_equal(oracle.roi_align_fwd(g["feat"][:1], rois), g["out_listform"])
    lv = oracle.level_map(g["ms_boxes"])
    assert np.array_equal(lv, g["ms_levels"])
    assert np.array_equal(oracle.level_map(g["lm_boxes"]), g["lm_levels"])


def test_multiscale_roi_align(golden, oracle, synth):
    g = golden("roi_align")
    feats = [synth.make_features(1, 8, 32 >> i, 40 >> i, seed=40 + i) for i in range(4)]
    rois = np.concatenate([np.zeros((40, 1), np.float32), g["ms_boxes"]], axis=1)
    out = oracle.multiscale_roi_align_fwd(feats, [0.25, 0.125, 0.0625, 0.03125], rois, g["ms_levels"])
    assert np.array_equal(out, g["ms_out"])


def test_paste_bit_exact(golden, oracle, synth):
    g = golden("paste")
    m = oracle.paste_masks(g["probs"], g["boxes"], 64, 80)
    assert np.array_equal(m, g["masks"])
    assert set(np.unique(m)) <= {0, 255}
    probs = synth.make_mask_probs(12, 28, seed=int(g["full_seed_probs"]))
    boxes = synth.make_det_boxes(12, int(g["full_seed_boxes"]))
    full = oracle.paste_masks(probs, boxes, 520, 704)
    assert np.array_equal(np.packbits(full > 0), g["full_masks_bits"])


def test_pipeline_chain(golden, oracle, synth):
    """The oracle chained like forward_inference (src/custom_maskrcnn.py:164-207) reproduces the
    reference's chained outputs."""
    g = golden("pipeline")
    base = oracle.base_anchors()
    for b in range(2):
        boxes, scores, _, keep = _nms_after_select(oracle, g["obj"][b], base, 200, 0.3, 0.4, 60, 10, 96, 128)
        props, ps = boxes[keep], scores[keep]
        assert np.array_equal(props, g[f"props_{b}"])
        rois = np.concatenate([np.zeros((len(props), 1), np.float32), props], axis=1)
        rf = oracle.roi_align_fwd(g["feat"][b:b + 1], rois)
        assert np.array_equal(rf, g[f"roi_feat_{b}"])
        bs = synth.make_box_scores((60,), 63 + b)[: len(props)]
        keep2 = oracle.nms(props, bs, 0.5, score_thresh=0.4, use_score_thresh=True)
        assert np.array_equal(props[keep2], g[f"det_boxes_{b}"])
        assert np.array_equal(bs[keep2], g[f"det_scores_{b}"])
        probs = synth.make_mask_probs(60, 28, 65 + b)[: len(keep2)]
        masks = oracle.paste_masks(probs, props[keep2], 96, 128)
        assert np.array_equal(masks, g[f"det_masks_{b}"])


def test_match_rows_golden(oracle, golden):
    """the oracle's box_iou / row max and mask targets against the reference's
    own outputs (torchvision.ops.box_iou(...).max(dim=1), extract_mask_target) — tests/golden/make_golden.py:gen_match."""
    g = golden("match")
    iou, mx, am = oracle.box_iou(g["boxes"], g["gt"])
    assert np.array_equal(iou, g["iou"], equal_nan=True)            # same fp32 operations: bit-exact, NaN for 0/0
    assert np.array_equal(mx, g["max_iou"], equal_nan=True)
    assert np.array_equal(am, g["argmax"])
    assert np.isnan(g["max_iou"]).any() and (g["argmax"] == 2).any()   # the fixture exercises NaN rows and the tie
    _, mx2, am2 = oracle.box_iou(g["boxes"], g["gt"], want_matrix=False)
    assert np.array_equal(mx2, mx, equal_nan=True) and np.array_equal(am2, am)
    # the threshold masks and sums that follow the row max (src/components/rpn.py:76-81, src/custom_maskrcnn.py:224-225),
    # against the same expressions evaluated by torch on the reference's own row max
    import torch
    tmx = torch.from_numpy(g["max_iou"])
    for pos_thr, neg_thr in ((0.5, 0.3), (0.4, None)):
        mx3, am3, pos, neg, cnt = oracle.match_boxes(g["boxes"], g["gt"], pos_thr, neg_thr)
        t_pos, t_neg = tmx >= pos_thr, tmx < (pos_thr if neg_thr is None else neg_thr)
        assert np.array_equal(pos, t_pos.numpy()) and np.array_equal(neg, t_neg.numpy())
        assert cnt.tolist() == [int(t_pos.sum().item()), int(t_neg.sum().item())]
        assert not (pos | neg)[np.isnan(mx3)].any()                     # NaN rows are in neither mask
        assert np.array_equal(mx3, mx, equal_nan=True) and np.array_equal(am3, am)
    tg = oracle.mask_targets(g["masks"], g["t_boxes"], g["t_index"], 28)
    assert tg.shape == g["targets"].shape
    # ATen's CPU bilinear kernel is FMA-contracted differently per build: values, not bits
    np.testing.assert_allclose(tg, g["targets"], rtol=0, atol=1e-6)


def _stitch_results(g):
    th, tw = [int(v) for v in g["tile_hw"]]
    n = int(g["n_per_tile"])
    res = []
    for t in range(25):
        masks = (np.unpackbits(g[f"t{t}_masks_bits"])[: n * th * tw].reshape(n, th, tw) * 255).astype(np.uint8)
        res.append(dict(tile_num=t, boxes=g[f"t{t}_boxes"], scores=g[f"t{t}_scores"], masks=masks))
    return res


def test_tail_and_stitch_rows_golden(oracle, golden):
    """the oracle's mask-head tail against the reference head's own interpolate + sigmoid, and its
    pixel counts driving the reference's tile filter (visualize.py:174-257) to the reference's kept set."""
    g = golden("tail_stitch")
    np.testing.assert_allclose(oracle.mask_tail(g["logits14"], 28, 1), g["probs28"], rtol=0, atol=1e-6)
    np.testing.assert_allclose(oracle.mask_tail(g["logits28"], 28, 1), g["probs28_same"], rtol=0, atol=1e-6)
    from livecell_instance_segmentation_b200 import stitch
    assert [len(stitch.get_valid_mini_tiles_for_tile(t)) for t in range(25)] == g["valid_mini_tiles"].tolist()
    # host logic of stitch.py with the oracle's counts in place of the kernel's
    mtw, mth = 704 // 7, 520 // 7
    kept, processed = [], set()
    for r in _stitch_results(g):
        t = r["tile_num"]
        cs, rs = stitch.get_tile_position_in_grid(t)
        ox, oy = cs * mtw, rs * mth
        new = [mt for mt in stitch.get_valid_mini_tiles_for_tile(t) if mt not in processed]
        if not new:
            continue
        k = r["scores"] > 0.5
        masks, boxes, scores = r["masks"][k], r["boxes"][k], r["scores"][k]
        h, w = masks.shape[1:]
        rects = [[max(0, mc * mtw - ox), max(0, mr * mth - oy), min(w, mc * mtw + mtw - ox), min(h, mr * mth + mth - oy)] for mc, mr in new]
        live = [q[0] < q[2] and q[1] < q[3] for q in rects]
        rects = [q if ok else [0, 0, 0, 0] for q, ok in zip(rects, live)]
        n, R = len(boxes), len(rects)
        total, inreg = oracle.mask_region_counts(masks, np.tile(np.array(rects, np.int32), (n, 1)), np.arange(0, (n + 1) * R, R))
        inreg = inreg.reshape(n, R)
        for i in range(n):
            frac = sum((int(inreg[i, q]) / int(total[i])) if (live[q] and total[i] > 0) else 0.0 for q in range(R))
            if frac > 0.4:
                kept.append((t, frac, float(scores[i]), [boxes[i][0] + ox, boxes[i][1] + oy, boxes[i][2] + ox, boxes[i][3] + oy]))
        processed.update(new)
    assert [k[0] for k in kept] == g["kept_tile"].tolist()
    assert np.array_equal(np.array([k[1] for k in kept]), g["kept_fraction"])          # float64, same summation order
    assert np.array_equal(np.array([k[2] for k in kept]), g["kept_score"])
    np.testing.assert_allclose(np.array([k[3] for k in kept]), g["kept_box"], rtol=0, atol=1e-4)


# ---- a12 / a13 (P2): torchvision's own RegionProposalNetwork and paste_masks_in_image ------------------------------------------
def _tv_case(tag):
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("make_golden_cases", os.path.join(os.path.dirname(__file__), "golden", "tv_cases.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.tv_rpn_case(tag)


@pytest.mark.parametrize("tag", ["trick", "vanilla"])
def test_torchvision_rpn_filter_proposals(oracle, golden, tag):
    """tests/golden/tv_rpn.npz = outputs of torchvision's RegionProposalNetwork.forward (eval) with torchvision's
    AnchorGenerator, BoxCoder and batched_nms on seeded objectness / delta maps (make_golden.py:gen_tv_rpn).  The oracle's
    restatement must reproduce: the anchors (sha256 of all of them), the per-level top-n index lists bit-exactly, and the final
    proposals — same boxes (decode: 1e-5 relative, exp), scores 1e-6, in the same order."""
    import hashlib
    g = golden("tv_rpn")
    c = _tv_case(tag)
    L = len(c["sizes"])
    bases = [oracle.tv_base_anchors(c["sizes"][l], c["ratios"]) for l in range(L)]
    assert np.array_equal(np.stack(bases), g[f"{tag}_base_anchors"])
    anchors = np.concatenate([oracle.anchors(h, w, c["strides"][l], bases[l]) for l, (h, w) in enumerate(c["shapes"])])
    assert np.array_equal(np.frombuffer(hashlib.sha256(anchors.tobytes()).digest(), np.uint8), g[f"{tag}_anchors_sha256"])
    trick = bool(g[f"{tag}_uses_coordinate_trick"])
    for b in range(c["B"]):
        boxes, scores, level, top = oracle.tv_rpn_filter_proposals(
            [o[b] for o in c["obj"]], [d[b] for d in c["deltas"]], bases, c["strides"], c["H"], c["W"], k=c["k"], post_n=c["post"],
            nms_thresh=c["nms"], score_thresh=c["score_thresh"], min_size=c["min_size"], coordinate_trick=trick)
        off, want_idx = 0, g[f"{tag}_top_n_idx"][b]
        pos = 0
        for l, (h, w) in enumerate(c["shapes"]):
            n = len(top[l])
            assert np.array_equal(top[l] + off, want_idx[pos:pos + n]), (tag, b, l)        # bit-exact top-n indices
            pos += n
            off += c["A"] * h * w
        assert pos == want_idx.size
        wb, ws = g[f"{tag}_boxes_{b}"], g[f"{tag}_scores_{b}"]
        assert len(boxes) == len(wb), (tag, b, len(boxes), len(wb))
        np.testing.assert_allclose(scores, ws, rtol=0, atol=1e-6)
        np.testing.assert_allclose(boxes, wb, rtol=1e-5, atol=1e-4)


def test_torchvision_paste_masks(oracle, golden):
    """tests/golden/tv_paste.npz = torchvision.models.detection.roi_heads.paste_masks_in_image(padding=1) on CPU."""
    g = golden("tv_paste")
    for tag in ("a", "b"):
        H, W = [int(v) for v in g[f"{tag}_size"]]
        out = oracle.paste_masks_tv(g[f"{tag}_probs"], g[f"{tag}_boxes"], H, W, padding=1)
        want = g[f"{tag}_out"]
        assert np.array_equal(out != 0, want != 0)                   # the same pixels are written
        np.testing.assert_allclose(out, want, rtol=0, atol=1e-6)     # ATen's CPU bilinear is FMA-contracted per build


def test_roi_align_c64_reference_fixture(oracle, golden, synth):
    """The 64-channel reference-run fixture (fast-path shape): the oracle reproduces torchvision's compiled CPU op bit for bit,
    forward and backward."""
    g = golden("roi_align_c64")
    s_feat, s_rois, s_g = [int(v) for v in g["seeds"]]
    feat = synth.make_features(1, 64, 40, 48, seed=s_feat)
    assert np.array_equal(g["rois"], synth.make_rois(16, s_rois, img_h=160, img_w=192, edge_cases=True))
    assert np.array_equal(oracle.roi_align_fwd(feat, g["rois"], 7, 7, 0.25, 2, False), g["out"])
    gout = np.random.RandomState(s_g).standard_normal((16, 64, 7, 7)).astype(np.float32)
    gin = oracle.roi_align_bwd(gout, g["rois"], feat.shape, 0.25, 2, False)
    np.testing.assert_allclose(gin, g["gin"], rtol=0, atol=1e-5 * np.abs(g["gin"]).max())

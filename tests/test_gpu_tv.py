"""GPU parity of the torchvision-semantics (P2) rows: a12 — RegionProposalNetwork's proposal stage (per-level top-k on
logits, decode, filters, batched_nms across levels, post-NMS top-n) and a13's torchvision paste — against fixtures generated
by torchvision's own code (tests/golden/make_golden.py:gen_tv_rpn / gen_tv_paste) and against the oracle."""
import importlib.util
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _case(tag):
    spec = importlib.util.spec_from_file_location("tv_cases", os.path.join(os.path.dirname(__file__), "golden", "tv_cases.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.tv_rpn_case(tag)


@pytest.mark.parametrize("tag", ["trick", "vanilla"])
def test_rpn_filter_proposals_matches_torchvision(golden, tag):
    from gpu_util import N, T
    from livecell_instance_segmentation_b200 import ops, tv_rpn
    g = golden("tv_rpn")
    c = _case(tag)
    L, B = len(c["sizes"]), c["B"]
    bases = [tv_rpn.base_anchors(c["sizes"][l], c["ratios"]) for l in range(L)]
    assert np.array_equal(np.stack([b.numpy() for b in bases]), g[f"{tag}_base_anchors"])
    # anchors of every level through the anchors kernel with torchvision's (rounded, h/w-ratio) base anchors
    import hashlib
    anc = torch.cat([ops.anchors(h, w, c["strides"][l], bases[l], "cuda:0") for l, (h, w) in enumerate(c["shapes"])])
    assert np.array_equal(np.frombuffer(hashlib.sha256(N(anc).tobytes()).digest(), np.uint8), g[f"{tag}_anchors_sha256"])
    objs, dls = [T(o) for o in c["obj"]], [T(d) for d in c["deltas"]]
    # (1) _get_top_n_idx: per-level top-k on the logits, no filter -> bit-exact index lists
    k = c["k"]
    _, _, idx, cnt = ops.rpn_select(objs, k=min(k, max(o[0].numel() for o in objs)), img_size=(c["H"], c["W"]), score_thresh=-1.0,
                                    min_size=-1.0, strides=c["strides"], base=bases, deltas=dls, score_strict=False, topk_on_sigmoid=False)
    idx, cnt = N(idx), N(cnt)
    for b in range(B):
        off, pos, want = 0, 0, g[f"{tag}_top_n_idx"][b]
        for l, (h, w) in enumerate(c["shapes"]):
            n = min(k, c["A"] * h * w)
            assert cnt[b, l] == n
            assert np.array_equal(idx[b, l, :n] + off, want[pos:pos + n]), (tag, b, l)
            pos += n
            off += c["A"] * h * w
    # (2) the whole stage, torchvision's CPU choice of batched_nms variant forced (the fixture was generated on CPU)
    trick = bool(g[f"{tag}_uses_coordinate_trick"])
    boxes, scores, counts, level = tv_rpn.filter_proposals(
        objs, dls, (c["H"], c["W"]), sizes=c["sizes"], aspect_ratios=c["ratios"], pre_nms_top_n=k, post_nms_top_n=c["post"],
        nms_thresh=c["nms"], score_thresh=c["score_thresh"], min_size=c["min_size"], coordinate_trick=trick, cpu_nms_threshold=True)
    boxes, scores, counts = N(boxes), N(scores), N(counts)
    for b in range(B):
        wb, ws = g[f"{tag}_boxes_{b}"], g[f"{tag}_scores_{b}"]
        n = int(counts[b])
        assert n == len(wb), (tag, b, n, len(wb))
        np.testing.assert_allclose(scores[b, :n], ws, rtol=0, atol=1e-6)         # same proposals in the same order
        np.testing.assert_allclose(boxes[b, :n], wb, rtol=1e-5, atol=1e-4)
    # the other batched_nms variant gives the same set on this data (no IoU within rounding of the threshold)
    b2, s2, c2, _ = tv_rpn.filter_proposals(
        objs, dls, (c["H"], c["W"]), sizes=c["sizes"], aspect_ratios=c["ratios"], pre_nms_top_n=k, post_nms_top_n=c["post"],
        nms_thresh=c["nms"], score_thresh=c["score_thresh"], min_size=c["min_size"], coordinate_trick=not trick, cpu_nms_threshold=True)
    assert np.array_equal(N(c2), counts) and np.array_equal(N(s2), scores)


def test_rpn_concat_levels_against_numpy():
    from gpu_util import N, T
    from livecell_instance_segmentation_b200 import ops
    rng = np.random.RandomState(3)
    B, L, k = 3, 5, 37
    boxes = rng.uniform(0, 300, size=(B, L, k, 4)).astype(np.float32)
    scores = rng.rand(B, L, k).astype(np.float32)
    counts = rng.randint(0, k + 1, size=(B, L)).astype(np.int32)
    counts[0] = 0                                                    # an image without survivors
    counts[1, 2] = k
    for trick in (True, False):
        cb, nb, cs, cl, cc = [N(t) for t in ops.rpn_concat_levels(T(boxes), T(scores), T(counts), trick)]
        for b in range(B):
            wb = np.concatenate([boxes[b, l, :counts[b, l]] for l in range(L)])
            ws = np.concatenate([scores[b, l, :counts[b, l]] for l in range(L)])
            wl = np.concatenate([np.full(counts[b, l], l, np.int32) for l in range(L)])
            n = len(ws)
            assert cc[b] == n
            assert np.array_equal(cb[b, :n], wb) and np.array_equal(cs[b, :n], ws) and np.array_equal(cl[b, :n], wl)
            assert (cl[b, n:] == -1).all() and not cb[b, n:].any()
            if n:
                shift = np.float32(wb.max()) + np.float32(1) if trick else np.float32(0)
                want = (wb + (wl.astype(np.float32) * shift).astype(np.float32)[:, None]).astype(np.float32)
                assert np.array_equal(nb[b, :n], want)


def test_paste_masks_torchvision_variant(golden, oracle, synth):
    from gpu_util import N, T
    from livecell_instance_segmentation_b200 import ops
    g = golden("tv_paste")
    for tag in ("a", "b"):
        H, W = [int(v) for v in g[f"{tag}_size"]]
        out = N(ops.paste_masks_tv(T(g[f"{tag}_probs"]), T(g[f"{tag}_boxes"]), H, W, padding=1))
        assert out.shape == (len(g[f"{tag}_boxes"]), 1, H, W) and out.dtype == np.float32
        want = g[f"{tag}_out"]
        assert np.array_equal(out[:, 0] != 0, want != 0)
        np.testing.assert_allclose(out[:, 0], want, rtol=0, atol=1e-6)
    # full frames, odd width (scalar tail stores), valid mask, vs the oracle: bit-exact (same FMA placement)
    for H, W, n in ((520, 704, 40), (37, 53, 9)):
        probs = synth.make_mask_probs(n, 28, 7 + H)
        rois = synth.make_rois(n, 11 + W, img_h=H, img_w=W, mode="fpn", edge_cases=n >= 8)[:, 1:]
        valid = np.ones(n, np.uint8)
        valid[n // 2] = 0
        buf = torch.full((n, 1, H, W), 7.0, device="cuda:0")
        out = N(ops.paste_masks_tv(T(probs), T(rois), H, W, padding=1, valid=T(valid), out=buf))
        want = oracle.paste_masks_tv(probs, rois, H, W, padding=1)
        keep = valid.astype(bool)
        assert np.array_equal(out[keep, 0], want[keep])
        assert (out[~keep] == 7.0).all()                              # skipped frames are left untouched

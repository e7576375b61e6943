#!/usr/bin/env python
"""Generate the golden fixtures in tests/golden/ by EXECUTING the reference on CPU.

Runs only in the authoring container (needs /root/reference and the installed torchvision CPU ops);
the GPU box never has /root/reference, it only reads the committed .npz files.

    python tests/golden/make_golden.py            # rewrites tests/golden/*.npz

Every array saved here is an input to, or an output of, a reference function (cited per fixture).
Inputs come from livecell_instance_segmentation_b200.synth with the seeds recorded in the file, so
large inputs are regenerated from the seed instead of being stored.
"""
import hashlib
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
sys.path.insert(0, HERE)

from livecell_instance_segmentation_b200 import synth  # noqa: E402

# --- the reference's own code (unmodified, imported from the read-only checkout) ---------------
from src.components.anchor_generator import AnchorGenerator  # noqa: E402
from src.utils.box_utils import clip_boxes_to_image, filter_small_boxes, encode_boxes  # noqa: E402
from src.utils.mask_utils import paste_masks_in_image, extract_mask_target  # noqa: E402
from src.utils.proposal_utils import generate_inference_proposals, generate_training_proposals  # noqa: E402
from torchvision.ops import RoIAlign, nms, MultiScaleRoIAlign, box_iou  # noqa: E402  (what custom_maskrcnn.py:5 imports)
from torchvision.ops.poolers import LevelMapper  # noqa: E402
from torchvision.models.detection._utils import BoxCoder  # noqa: E402

torch.set_num_threads(1)
T = torch.from_numpy


def save(name, **arrays):
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **arrays)
    print(f"{name}.npz  {os.path.getsize(path) / 1024:.1f} KiB")


def sha(a):
    return np.frombuffer(hashlib.sha256(np.ascontiguousarray(a).tobytes()).digest(), dtype=np.uint8)


def ref_topk_indices(obj, k):
    """The top-k index list the reference computes internally (proposal_utils.py:16-19 / :38-41)."""
    objectness = torch.sigmoid(T(obj)).permute(1, 2, 0).reshape(-1)
    s, i = torch.topk(objectness, min(k, objectness.numel()))
    return s.numpy(), i.numpy()


# ---------------------------------------------------------------------------------------------
def gen_anchors():
    g = AnchorGenerator()
    small = g.generate_anchors((5, 7), 4, "cpu").numpy()
    full = g.generate_anchors((130, 176), 4, "cpu").numpy()          # C1 level-0 map
    tile = g.generate_anchors((56, 75), 4, "cpu").numpy()            # real 300x222 tiles
    rows = np.array([0, 1, 8, 9, 1583, 102959, 205919])
    save("anchors", small=small, full_sha256=sha(full), full_rows=rows, full_vals=full[rows],
         tile_sha256=sha(tile), sizes=np.array(g.sizes, np.float64), ratios=np.array(g.aspect_ratios, np.float64))


def gen_box_utils():
    rng = np.random.RandomState(7)
    boxes = rng.uniform(-60, 760, size=(64, 4)).astype(np.float32)
    clipped = clip_boxes_to_image(T(boxes.copy()), (520, 704)).numpy()
    keep10 = filter_small_boxes(T(clipped), 10).numpy()
    keep5 = filter_small_boxes(T(clipped), 5).numpy()
    anc = synth.make_rois(64, 3)[:, 1:]
    gt = anc + rng.uniform(-6, 6, size=anc.shape).astype(np.float32)
    enc = encode_boxes(T(gt), T(anc)).numpy()
    coder = BoxCoder((1.0, 1.0, 1.0, 1.0))
    deltas = rng.normal(0, 0.3, size=(64, 4)).astype(np.float32)
    deltas[0, 2] = 9.0  # exercises the bbox_xform_clip clamp
    dec = coder.decode_single(T(deltas), T(anc)).numpy()
    coder2 = BoxCoder((10.0, 10.0, 5.0, 5.0))
    dec2 = coder2.decode_single(T(deltas), T(anc)).numpy()
    save("box_utils", boxes=boxes, clipped=clipped, keep10=keep10, keep5=keep5, anchors=anc, gt=gt, encoded=enc,
         deltas=deltas, decoded_w1=dec, decoded_w10=dec2)


def gen_proposals():
    g = AnchorGenerator()
    out = {}
    # small: 96x128 image, 24x32 map
    obj = synth.make_objectness(2, 9, 24, 32, n_cells=40, seed=11, k=100)
    anc = g.generate_anchors((24, 32), 4, "cpu")
    for b in range(2):
        p, s = generate_inference_proposals(T(obj[b]), anc, (96, 128), "cpu", num_pre_nms=100, score_threshold=0.3,
                                            nms_threshold=0.4, num_post_nms=30, min_box_size=10)
        out[f"small_inf_boxes_{b}"], out[f"small_inf_scores_{b}"] = p.numpy(), s.numpy()
        tp = generate_training_proposals(T(obj[b]), anc, (96, 128), "cpu", num_proposals=120, score_threshold=0.01,
                                         min_box_size=5)
        out[f"small_train_boxes_{b}"] = tp.numpy()
        ts, ti = ref_topk_indices(obj[b], 100)
        out[f"small_topk_scores_{b}"], out[f"small_topk_index_{b}"] = ts, ti
    out["small_obj"] = obj
    # C1: 704x520, reference defaults (250 -> 0.3 -> nms 0.4 -> 50)
    obj = synth.make_objectness(1, 9, 130, 176, n_cells=150, seed=21, k=250)
    anc = g.generate_anchors((130, 176), 4, "cpu")
    p, s = generate_inference_proposals(T(obj[0]), anc, (520, 704), "cpu")
    ts, ti = ref_topk_indices(obj[0], 250)
    out.update(c1_seed=np.array(21), c1_inf_boxes=p.numpy(), c1_inf_scores=s.numpy(), c1_topk_index=ti, c1_topk_scores=ts)
    # C2: 256x256 tile, training defaults (500, 0.01, 5)
    obj = synth.make_objectness(1, 9, 64, 64, n_cells=20, seed=22, k=500)
    anc2 = g.generate_anchors((64, 64), 4, "cpu")
    tp = generate_training_proposals(T(obj[0]), anc2, (256, 256), "cpu")
    ts, ti = ref_topk_indices(obj[0], 500)
    out.update(c2_seed=np.array(22), c2_train_boxes=tp.numpy(), c2_topk_index=ti)
    # C3: crowded, 2000 cells, k=2000, post 1000
    obj = synth.make_objectness(1, 9, 130, 176, n_cells=2000, seed=23, k=2000)
    p, s = generate_inference_proposals(T(obj[0]), anc, (520, 704), "cpu", num_pre_nms=2000, num_post_nms=1000)
    ts, ti = ref_topk_indices(obj[0], 2000)
    out.update(c3_seed=np.array(23), c3_inf_boxes=p.numpy(), c3_inf_scores=s.numpy(), c3_topk_index=ti)
    save("proposals", **out)


def gen_nms():
    out = {}
    rng = np.random.RandomState(5)
    for n in (250, 2000):
        r = synth.make_rois(n, 100 + n)[:, 1:]
        sc = rng.permutation(n).astype(np.float32) / n       # unique scores
        out[f"boxes_{n}"], out[f"scores_{n}"] = r, sc
        for thr in (0.4, 0.5, 0.7):
            out[f"keep_{n}_{int(thr * 10)}"] = nms(T(r), T(sc), thr).numpy()
    # known-answer probes (SURVEY §8c "oracle facts")
    b = np.array([[0, 0, 10, 10], [0, 0, 10, 10], [20, 20, 30, 30], [0, 0, 10, 10], [21, 21, 31, 31]], np.float32)
    s = np.array([0.5, 0.9, 0.7, 0.9, 0.7], np.float32)      # ties -> lower index first
    out["tie_boxes"], out["tie_scores"], out["tie_keep"] = b, s, nms(T(b), T(s), 0.5).numpy()
    b = np.array([[0, 0, 10, 10], [0, 0, 4, 10], [0, 0, 10, 5]], np.float32)   # IoU exactly 0.4f and 0.5
    s = np.array([0.9, 0.8, 0.7], np.float32)
    out["eq_boxes"], out["eq_scores"] = b, s
    out["eq_keep_04"], out["eq_keep_05"] = nms(T(b), T(s), 0.4).numpy(), nms(T(b), T(s), 0.5).numpy()
    b = np.array([[5, 5, 5, 5], [5, 5, 5, 5], [5, 5, 5, 9], [0, 0, 10, 10]], np.float32)   # zero area: NaN IoU
    s = np.array([0.9, 0.8, 0.7, 0.6], np.float32)
    out["zero_boxes"], out["zero_scores"], out["zero_keep"] = b, s, nms(T(b), T(s), 0.4).numpy()
    b = np.array([[0, 0, 10, 10], [1, 1, 11, 11], [50, 50, 60, 60]], np.float32)
    s = np.array([0.5, np.nan, 0.7], np.float32)             # NaN score sorts first
    out["nan_boxes"], out["nan_scores"], out["nan_keep"] = b, s, nms(T(b), T(s), 0.4).numpy()
    save("nms", **out)


def gen_roi_align():
    out = {}
    feat = synth.make_features(2, 8, 20, 24, seed=31)
    rois = synth.make_rois(24, 32, img_h=80, img_w=96, batch=2, edge_cases=True)
    out["feat"], out["rois"] = feat, rois
    rng = np.random.RandomState(33)
    for tag, (P, sr, al) in {"p7": (7, 2, False), "p14": (14, 2, False), "p7a": (7, 2, True), "p7ad": (7, 0, False)}.items():
        op = RoIAlign(output_size=(P, P), spatial_scale=0.25, sampling_ratio=sr, aligned=al)
        f = T(feat).clone().requires_grad_(True)
        y = op(f, T(rois))
        g = rng.standard_normal(tuple(y.shape)).astype(np.float32)
        y.backward(T(g))
        out[f"out_{tag}"], out[f"gout_{tag}"], out[f"gin_{tag}"] = y.detach().numpy(), g, f.grad.numpy()
    # the reference call form: feature_map[b:b+1], [proposals]  (src/custom_maskrcnn.py:177)
    op = RoIAlign(output_size=(7, 7), spatial_scale=1.0 / 4.0, sampling_ratio=2)
    out["out_listform"] = op(T(feat[:1]), [T(rois[:, 1:])]).numpy()
    # multi-level (P2): MultiScaleRoIAlign over 4 levels of a 128x160 image
    feats = {str(i): T(synth.make_features(1, 8, 32 >> i, 40 >> i, seed=40 + i)) for i in range(4)}
    boxes = synth.make_rois(40, 41, img_h=128, img_w=160, mode="fpn")[:, 1:]
    ms = MultiScaleRoIAlign(["0", "1", "2", "3"], 7, 2)
    out["ms_out"] = ms(feats, [T(boxes)], [(128, 160)]).numpy()
    out["ms_boxes"] = boxes
    out["ms_levels"] = LevelMapper(2, 5)([T(boxes)]).numpy()
    big = synth.make_rois(4096, 42, mode="fpn")[:, 1:]
    out["lm_boxes"], out["lm_levels"] = big, LevelMapper(2, 5)([T(big)]).numpy()
    save("roi_align", **out)
    # C = 64: a reference-run fixture that reaches the warp-item / staged fast paths (C % 64 == 0, channels_last),
    # forward and backward; inputs are regenerated from the seeds, only torchvision's outputs are stored
    feat = synth.make_features(1, 64, 40, 48, seed=51)
    rois = synth.make_rois(16, 52, img_h=160, img_w=192, edge_cases=True)
    op = RoIAlign(output_size=(7, 7), spatial_scale=0.25, sampling_ratio=2)
    f = T(feat).clone().requires_grad_(True)
    y = op(f, T(rois))
    g = np.random.RandomState(53).standard_normal(tuple(y.shape)).astype(np.float32)
    y.backward(T(g))
    save("roi_align_c64", out=y.detach().numpy(), gin=f.grad.numpy(), seeds=np.array([51, 52, 53]), rois=rois)


def gen_paste():
    out = {}
    probs = synth.make_mask_probs(8, 28, seed=51)
    boxes = synth.make_det_boxes(8, 52, img_h=64, img_w=80, lo=6, hi=40, edge_cases=True)
    out["probs"], out["boxes"] = probs, boxes
    out["masks"] = paste_masks_in_image(T(probs), T(boxes), (64, 80), threshold=0.5).numpy()
    # full frame 704x520, 12 detections, stored bit-packed (values are {0,255})
    probs = synth.make_mask_probs(12, 28, seed=53)
    boxes = synth.make_det_boxes(12, 54)
    m = paste_masks_in_image(T(probs), T(boxes), (520, 704)).numpy()
    assert set(np.unique(m)) <= {0, 255}
    out["full_seed_probs"], out["full_seed_boxes"] = np.array(53), np.array(54)
    out["full_masks_bits"] = np.packbits(m > 0)
    save("paste", **out)


def gen_pipeline():
    """Region pipeline chained exactly as forward_inference does (src/custom_maskrcnn.py:164-207),
    with the PyTorch heads replaced by seeded synthetic scores / mask probabilities."""
    g = AnchorGenerator()
    B, h, w, H, W = 2, 24, 32, 96, 128
    obj = synth.make_objectness(B, 9, h, w, n_cells=60, seed=61, k=200)
    feat = synth.make_features(B, 16, h, w, seed=62)
    anc = g.generate_anchors((h, w), 4, "cpu")
    roi_align = RoIAlign(output_size=(7, 7), spatial_scale=0.25, sampling_ratio=2)
    out = dict(obj=obj, feat=feat)
    for b in range(B):
        props, sc = generate_inference_proposals(T(obj[b]), anc, (H, W), "cpu", num_pre_nms=200, num_post_nms=60)
        rf = roi_align(T(feat[b:b + 1]), [props])
        box_scores = T(synth.make_box_scores((60,), 63 + b))[: len(props)]
        keep_scores = box_scores > 0.4
        fb, fs = props[keep_scores], box_scores[keep_scores]
        keep = nms(fb, fs, 0.5)
        fb, fs = fb[keep], fs[keep]
        probs = T(synth.make_mask_probs(60, 28, 65 + b))[: len(fb)]
        masks = paste_masks_in_image(probs, fb, (H, W))
        out[f"props_{b}"], out[f"prop_scores_{b}"], out[f"roi_feat_{b}"] = props.numpy(), sc.numpy(), rf.numpy()
        out[f"det_boxes_{b}"], out[f"det_scores_{b}"], out[f"det_masks_{b}"] = fb.numpy(), fs.numpy(), masks.numpy()
    save("pipeline", **out)


def gen_match():
    """box_iou(...).max(dim=1) (src/components/rpn.py:72-73, src/custom_maskrcnn.py:221-222)
    and extract_mask_target (src/utils/mask_utils.py:6-46), executed on CPU."""
    rng = np.random.RandomState(91)
    g = AnchorGenerator()
    anc = g.generate_anchors((16, 20), 4, "cpu")                       # 2880 anchors of a 64x80 tile
    gt = synth.make_det_boxes(12, 92, img_h=64, img_w=80, lo=8, hi=40)
    gt[3] = gt[2]                                                      # duplicate box: argmax tie -> first index
    gt[7] = [10.0, 10.0, 10.0, 10.0]                                   # zero-area gt
    boxes = torch.cat([anc, T(np.array([[10.0, 10.0, 10.0, 10.0], [5.0, 5.0, 4.0, 4.0]], np.float32))])   # 0/0 and inverted
    ious = box_iou(boxes, T(gt))
    mx, am = ious.max(dim=1)
    out = dict(boxes=boxes.numpy(), gt=gt, iou=ious.numpy(), max_iou=mx.numpy(), argmax=am.numpy())
    # mask targets: rectangular + elliptical uint8 masks on a 64x80 tile, boxes incl. out-of-frame and sub-pixel ones
    H, W, G = 64, 80, 6
    yy, xx = np.mgrid[0:H, 0:W]
    masks = np.zeros((G, H, W), np.uint8)
    mb = synth.make_det_boxes(G, 93, img_h=H, img_w=W, lo=10, hi=36)
    for i in range(G):
        x1, y1, x2, y2 = mb[i]
        cx, cy, rx, ry = (x1 + x2) / 2, (y1 + y2) / 2, max((x2 - x1) / 2, 1), max((y2 - y1) / 2, 1)
        masks[i] = (((xx - cx) / rx) ** 2 + ((yy - cy) / ry) ** 2 <= 1.0) if i % 2 else ((xx >= x1) & (xx < x2) & (yy >= y1) & (yy < y2))
    tb = np.concatenate([mb + rng.uniform(-3, 3, size=mb.shape).astype(np.float32), mb[:4] * 0.5,
                         np.array([[-9.5, -4.0, 20.3, 30.9], [70.2, 50.1, 95.0, 80.0], [30.4, 20.2, 30.9, 20.7],
                                   [79.6, 63.5, 85.0, 70.0]], np.float32)]).astype(np.float32)
    idx = rng.randint(0, G, size=len(tb)).astype(np.int64)
    tg = torch.stack([extract_mask_target(T(masks[i]), T(b), 28) for i, b in zip(idx, tb)])
    out.update(masks=masks, t_boxes=tb, t_index=idx, targets=tg.numpy())
    save("match", **out)


def _import_reference_visualize():
    """src/visualize.py needs matplotlib / pycocotools at import time (not installed here): stub them —
    the functions executed below (tile geometry + mask-area filter, visualize.py:100-257) use neither."""
    import types
    for name in ("matplotlib", "matplotlib.pyplot", "matplotlib.patches", "pycocotools", "pycocotools.mask"):
        if name not in sys.modules:
            sys.modules[name] = types.ModuleType(name)
    sys.modules["matplotlib"].pyplot = sys.modules["matplotlib.pyplot"]
    sys.modules["matplotlib"].patches = sys.modules["matplotlib.patches"]
    sys.modules["pycocotools"].mask = sys.modules["pycocotools.mask"]
    import importlib
    return importlib.import_module("src.visualize")


def gen_tail_stitch():
    """SURVEY §8(f) ranks 3-4.  (3) the tail of the reference mask head: CustomMaskHead.forward's final
    F.interpolate (src/components/mask_head.py:52-58) is executed through the reference module itself (the convs are
    bypassed by feeding mask_fcn_logits' output) + sigmoid(mask_logits[:, 1]) (src/custom_maskrcnn.py:273-274).
    (4) filter_detections_by_border_mini_tiles (src/visualize.py:174-257) on 25 synthetic tile predictions."""
    import torch.nn.functional as F
    from src.components.mask_head import CustomMaskHead
    rng = np.random.RandomState(97)
    head = CustomMaskHead()
    logits14 = rng.standard_normal((9, 2, 14, 14)).astype(np.float32) * 2
    # the head's own tail: everything before it replaced by identities so forward(x) == interpolate(x)
    ident = torch.nn.Identity()
    for name in ("conv1", "conv2", "conv3", "conv4", "deconv", "deconv_relu", "mask_fcn_logits"):
        setattr(head, name, ident)
    up = head(T(logits14))
    assert tuple(up.shape) == (9, 2, 28, 28)
    probs = torch.sigmoid(up[:, 1])
    out = dict(logits14=logits14, probs28=probs.numpy())
    logits28 = rng.standard_normal((3, 2, 28, 28)).astype(np.float32)
    out.update(logits28=logits28, probs28_same=torch.sigmoid(head(T(logits28))[:, 1]).numpy())

    viz = _import_reference_visualize()
    th, tw = 3 * (viz.IMG_HEIGHT // viz.N_MINI_ROWS), 3 * (viz.IMG_WIDTH // viz.N_MINI_COLS)     # 222 x 300 tiles
    results, packed = [], {}
    for t in range(viz.TOTAL_TILES):
        n = 14
        boxes = synth.make_det_boxes(n, 200 + t, img_h=th, img_w=tw, lo=14, hi=60)
        probs_t = synth.make_mask_probs(n, 28, 300 + t)
        masks = paste_masks_in_image(T(probs_t), T(boxes), (th, tw)).numpy()
        scores = rng.uniform(0.2, 1.0, size=n).astype(np.float32)
        results.append({"tile_num": t, "prediction": {"boxes": T(boxes), "scores": T(scores), "masks": T(masks)}})
        packed[f"t{t}_boxes"], packed[f"t{t}_scores"], packed[f"t{t}_masks_bits"] = boxes, scores, np.packbits(masks > 0)
    kept = viz.filter_detections_by_border_mini_tiles(list(reversed(results)), score_threshold=0.5, mask_threshold=0.4)
    assert len(kept) > 20
    out.update(packed)
    out.update(tile_hw=np.array([th, tw]), n_per_tile=np.array(14),
               kept_tile=np.array([d["tile_num"] for d in kept]), kept_box=np.array([d["box"] for d in kept], np.float64),
               kept_score=np.array([d["score"] for d in kept], np.float64),
               kept_fraction=np.array([d["area_fraction"] for d in kept], np.float64),
               kept_mask_sum=np.array([int(d["mask"].sum()) for d in kept]),
               valid_mini_tiles=np.array([len(viz.get_valid_mini_tiles_for_tile(t)) for t in range(viz.TOTAL_TILES)]))
    save("tail_stitch", **out)



from tv_cases import tv_rpn_case  # noqa: E402  (seeded inputs shared with the tests)


def gen_tv_rpn():
    """a12 — torchvision's own RegionProposalNetwork (TV:models/detection/rpn.py:231-297 filter_proposals / _get_top_n_idx,
    :339-367 forward) with torchvision's AnchorGenerator (TV:models/detection/anchor_utils.py:58-133: ratio = h/w, rounded
    base anchors), BoxCoder.decode and batched_nms, executed on CPU.  The head is a stand-in that returns the seeded
    objectness / delta maps, everything after it is torchvision's code."""
    from torchvision.models.detection.anchor_utils import AnchorGenerator as TVAnchorGenerator
    from torchvision.models.detection.image_list import ImageList
    from torchvision.models.detection.rpn import RegionProposalNetwork

    class Head(torch.nn.Module):
        def __init__(self, obj, deltas):
            super().__init__()
            self.obj, self.deltas = obj, deltas

        def forward(self, features):
            return [T(o) for o in self.obj], [T(d) for d in self.deltas]

    out = {}
    for tag in ("trick", "vanilla"):
        c = tv_rpn_case(tag)
        L, B = len(c["sizes"]), c["B"]
        ag = TVAnchorGenerator(sizes=c["sizes"], aspect_ratios=(c["ratios"],) * L)
        rpn = RegionProposalNetwork(ag, Head(c["obj"], c["deltas"]), fg_iou_thresh=0.7, bg_iou_thresh=0.3, batch_size_per_image=256,
                                    positive_fraction=0.5, pre_nms_top_n=dict(training=c["k"], testing=c["k"]),
                                    post_nms_top_n=dict(training=c["post"], testing=c["post"]), nms_thresh=c["nms"],
                                    score_thresh=c["score_thresh"])
        rpn.min_size = c["min_size"]
        rpn.eval()
        images = ImageList(torch.zeros((B, 3, c["H"], c["W"])), [(c["H"], c["W"])] * B)
        feats = {str(l): torch.zeros((B, 8, h, w)) for l, (h, w) in enumerate(c["shapes"])}
        captured = {}
        inner = rpn.filter_proposals

        def spy(proposals, objectness, image_shapes, num_anchors_per_level):
            captured["proposals"] = proposals.clone()
            captured["top_n_idx"] = rpn._get_top_n_idx(objectness.detach().reshape(B, -1), num_anchors_per_level)
            captured["n_per_level"] = list(num_anchors_per_level)
            return inner(proposals, objectness, image_shapes, num_anchors_per_level)

        rpn.filter_proposals = spy
        with torch.no_grad():
            boxes, _ = rpn(images, feats)
        # scores are not returned by forward(): recompute them exactly as filter_proposals does
        with torch.no_grad():
            fb, fs = inner(captured["proposals"], torch.cat([T(o).permute(0, 2, 3, 1).reshape(B, -1) for o in c["obj"]], dim=1).reshape(-1, 1),
                           images.image_sizes, captured["n_per_level"])
        anchors = ag(images, list(feats.values()))[0].numpy()
        numel = [int(min(c["k"], n)) for n in captured["n_per_level"]]
        out[f"{tag}_anchors_sha256"] = sha(anchors)
        out[f"{tag}_anchor_rows"] = anchors[[0, 1, 2, 3, anchors.shape[0] // 2, anchors.shape[0] - 1]]
        out[f"{tag}_base_anchors"] = np.stack([b.numpy() for b in ag.cell_anchors])
        out[f"{tag}_top_n_idx"] = captured["top_n_idx"].numpy().astype(np.int32)          # [B, sum_l min(k, n_l)], level offsets included
        out[f"{tag}_pre_nms_per_level"] = np.array(numel)
        out[f"{tag}_uses_coordinate_trick"] = np.array(sum(numel) * 4 <= 4000)
        for b in range(B):
            assert torch.equal(boxes[b], fb[b])
            s = fs[b].numpy()
            assert np.unique(s).size == s.size, "cross-level score tie in the fixture: change the seed"
            out[f"{tag}_boxes_{b}"], out[f"{tag}_scores_{b}"] = boxes[b].numpy(), s
        print(f"tv_rpn[{tag}]: {[len(b) for b in boxes]} proposals, coordinate trick: {bool(out[f'{tag}_uses_coordinate_trick'])}")
    save("tv_rpn", **out)


def gen_tv_paste():
    """a13, P2 variant — torchvision's paste_masks_in_image (TV:models/detection/roi_heads.py:405-501): boxes expanded by
    (M + 2) / M with the mask zero-padded by 1 px, integer box with +1 extents (TO_REMOVE = 1), bilinear resize of the
    PROBABILITIES (align_corners=False), float32 result [N, 1, H, W] (no threshold)."""
    from torchvision.models.detection.roi_heads import paste_masks_in_image as tv_paste
    rng = np.random.RandomState(77)
    out = {}
    for tag, (H, W, N_) in {"a": (96, 128, 24), "b": (222, 300, 12)}.items():
        probs = synth.make_mask_probs(N_, 28, 300 + H)
        w = rng.uniform(4, 70, N_)
        h = rng.uniform(4, 70, N_)
        x1 = rng.uniform(-10, W - 8, N_)
        y1 = rng.uniform(-10, H - 8, N_)
        boxes = np.stack([x1, y1, x1 + w, y1 + h], 1).astype(np.float32)
        boxes[0] = [10.2, 12.7, 10.9, 40.3]             # narrower than a pixel
        boxes[1] = [-20.0, -15.0, W + 30.0, H + 25.0]   # larger than the frame
        boxes[2] = [W - 3.5, H - 2.5, W + 20.0, H + 20.0]
        res = tv_paste(T(probs)[:, None], T(boxes), (H, W), padding=1)
        out[f"{tag}_probs"], out[f"{tag}_boxes"], out[f"{tag}_out"] = probs, boxes, res.numpy()[:, 0]
        out[f"{tag}_size"] = np.array([H, W])
    save("tv_paste", **out)


if __name__ == "__main__":
    if "--only-roi" in sys.argv:
        gen_roi_align()
        sys.exit(0)
    if "--only-tv" in sys.argv:
        gen_tv_rpn()
        gen_tv_paste()
        sys.exit(0)
    if "--only-match" in sys.argv:
        gen_match()
        sys.exit(0)
    if "--only-tail-stitch" in sys.argv:
        gen_tail_stitch()
        sys.exit(0)
    gen_anchors()
    gen_box_utils()
    gen_proposals()
    gen_nms()
    gen_roi_align()
    gen_paste()
    gen_pipeline()
    gen_match()
    gen_tail_stitch()
    gen_tv_rpn()
    gen_tv_paste()

"""CPU: host-side logic — shard arithmetic, config defaults mirroring the reference, the install()
hook's rebinding, synthetic-input determinism."""
import sys
import types

import numpy as np
import pytest


def test_shard_ranges_cover_everything():
    from livecell_instance_segmentation_b200.dist import max_shard, shard_range
    for n in (0, 1, 7, 64, 512, 513):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [e - s for s, e in spans]
            assert max(sizes) - min(sizes) <= 1 and max(sizes) == (max_shard(n, w) if n else 0)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def test_config_defaults_are_the_reference_kwargs():
    """src/utils/proposal_utils.py:33-36, src/custom_maskrcnn.py:48-50,185,192,292."""
    from livecell_instance_segmentation_b200.pipeline import RegionConfig
    c = RegionConfig()
    assert (c.pre_nms_top_n, c.rpn_score_thresh, c.rpn_nms_thresh, c.post_nms_top_n, c.min_box_size) == (250, 0.3, 0.4, 50, 10)
    assert (c.pooled_size, c.spatial_scale, c.sampling_ratio) == (7, 0.25, 2)
    assert (c.box_score_thresh, c.det_nms_thresh, c.mask_thresh, c.mask_size, c.stride) == (0.4, 0.5, 0.5, 28, 4)
    assert c.det_capacity == 50 and RegionConfig(max_detections=500).det_capacity == 500


def test_shim_signatures_mirror_the_reference():
    import inspect
    from livecell_instance_segmentation_b200.src.utils import proposal_utils as pu, mask_utils as mu, box_utils as bu
    from livecell_instance_segmentation_b200.src.components.anchor_generator import AnchorGenerator
    from livecell_instance_segmentation_b200.roi_align import RoIAlign, nms
    sig = inspect.signature(pu.generate_inference_proposals)
    assert list(sig.parameters) == ["cls_scores", "anchors", "image_size", "device", "num_pre_nms", "score_threshold",
                                    "nms_threshold", "num_post_nms", "min_box_size"]
    assert [p.default for p in list(sig.parameters.values())[4:]] == [250, 0.3, 0.4, 50, 10]
    sig = inspect.signature(pu.generate_training_proposals)
    assert [p.default for p in list(sig.parameters.values())[4:]] == [500, 0.01, 5]
    assert inspect.signature(pu.sample_proposals).parameters["num_samples"].default == 128
    assert list(inspect.signature(mu.paste_masks_in_image).parameters) == ["masks", "boxes", "image_size", "threshold"]
    assert list(inspect.signature(mu.extract_mask_target).parameters) == ["gt_mask", "box", "mask_size"]
    assert list(inspect.signature(mu.compute_mask_loss_from_gt).parameters) == ["mask_logits", "proposals", "targets", "device", "mask_size"]
    assert inspect.signature(mu.compute_mask_loss_from_gt).parameters["mask_size"].default == 28
    assert inspect.signature(bu.filter_small_boxes).parameters["min_size"].default == 1
    g = AnchorGenerator()
    assert (g.sizes, g.aspect_ratios, g.num_anchors_per_location) == ((32, 64, 128), (0.5, 1.0, 2.0), 9)
    r = RoIAlign(output_size=(7, 7), spatial_scale=0.25, sampling_ratio=2)
    assert len(list(r.parameters())) == 0 and len(list(r.buffers())) == 0 and r.aligned is False
    assert list(inspect.signature(nms).parameters) == ["boxes", "scores", "iou_threshold"]


def test_install_rebinds_reference_names(monkeypatch):
    """install() swaps the names the reference binds at import time (src/custom_maskrcnn.py:5,12-19)."""
    from livecell_instance_segmentation_b200 import install as inst, roi_align as ra
    fake = types.ModuleType("src.custom_maskrcnn")

    class CustomMaskRCNN:                              # stand-in with the patched method
        def _generate_masks(self, *a):
            return "reference"

    for name in ("AnchorGenerator", "RoIAlign", "nms", "generate_training_proposals", "generate_inference_proposals",
                 "sample_proposals"):
        setattr(fake, name, object())
    fake.CustomMaskRCNN = CustomMaskRCNN
    fake.box_iou = "untouched"
    monkeypatch.setitem(sys.modules, "src.custom_maskrcnn", fake)
    monkeypatch.setitem(sys.modules, "custom_maskrcnn", fake)
    done = inst.install(import_missing=False)
    assert fake.RoIAlign is ra.RoIAlign and fake.nms is ra.nms and fake.box_iou is ra.box_iou   # §8f rank 1: one-kernel IoU
    assert "CustomMaskRCNN._generate_masks" in done["src.custom_maskrcnn"]
    assert CustomMaskRCNN._generate_masks is inst._paste_method


def test_install_on_the_real_reference_if_present(monkeypatch):
    """In the authoring container the reference checkout exists: the hook must patch its real modules
    and the patched model must fail loudly (no CPU fallback) instead of silently using torchvision."""
    import os
    if not os.path.isdir("/root/reference/src"):
        pytest.skip("reference checkout not present (GPU box)")
    import torch
    monkeypatch.syspath_prepend("/root/reference")
    for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        monkeypatch.delitem(sys.modules, m)
    from livecell_instance_segmentation_b200 import install as inst, roi_align as ra
    from livecell_instance_segmentation_b200._lib import LcrError
    done = inst.install()
    import src.custom_maskrcnn as cm
    assert cm.RoIAlign is ra.RoIAlign and "generate_inference_proposals" in done["src.custom_maskrcnn"]
    model = cm.get_custom_model()
    assert isinstance(model.roi_align, ra.RoIAlign)
    assert cm.count_parameters(model)["roi_align"] == 0          # checkpoints interchange
    if not torch.cuda.is_available():
        model.eval()
        with pytest.raises(LcrError):
            model([torch.rand(3, 64, 64)])
    for m in [k for k in sys.modules if k == "src" or k.startswith("src.")]:
        monkeypatch.delitem(sys.modules, m, raising=False)


def test_synth_is_deterministic_and_tie_free(synth):
    a = synth.make_objectness(2, 9, 24, 32, n_cells=40, seed=11, k=100)
    b = synth.make_objectness(2, 9, 24, 32, n_cells=40, seed=11, k=100)
    assert np.array_equal(a, b)
    s = (1.0 / (1.0 + np.exp(-a[0].astype(np.float64)))).astype(np.float32).reshape(-1)
    top = np.sort(s)[::-1][:101]
    assert np.all(np.diff(top.astype(np.float64)) < 0)
    r = synth.make_rois(16, 3, edge_cases=True)
    assert r.shape == (16, 5) and r.dtype == np.float32
    p = synth.make_mask_probs(3, 28, 1)
    assert p.shape == (3, 28, 28) and 0 < p.min() and p.max() < 1

"""CPU side of the drop-in tests: the staged reference (baseline/_ref) imports and runs UNCHANGED through the drivers of
tests/dropin_cases.py (untouched vs untouched — the harness, the stubs for matplotlib/gradio/pycocotools and the
determinism assumptions are what is checked here; the untouched-vs-patched comparison needs the B200:
tests/test_gpu_dropin.py).  Skipped where baseline/_ref was not staged."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import dropin_cases as dc  # noqa: E402
import ref_harness  # noqa: E402

pytestmark = pytest.mark.skipif(not ref_harness.available(), reason="baseline/_ref not staged (tools/stage_reference.py)")


def test_manifest_matches_the_staged_files():
    import hashlib
    import json
    with open(os.path.join(ref_harness.REF, "MANIFEST.json")) as f:
        man = json.load(f)["files"]
    assert "src/custom_maskrcnn.py" in man and len(man) >= 18
    for rel, sha in man.items():
        with open(os.path.join(ref_harness.REF, rel), "rb") as f:
            assert hashlib.sha256(f.read()).hexdigest() == sha, rel + " was modified after staging"


def test_inference_driver_is_deterministic_on_cpu():
    a, ia = dc.run_inference("cpu", False, 96, 128, 2, seed=0, n_cells=20)
    b, ib = dc.run_inference("cpu", False, 96, 128, 2, seed=0, n_cells=20, state=ia["state"])
    assert ia["roi_align_type"].startswith("torchvision")
    assert dc.compare_predictions(a, b, score_atol=0.0) == 0
    assert sum(len(p["boxes"]) for p in a) > 0, "the randomly initialised model should produce detections"


def test_train_step_driver_has_positives_and_repeats_on_cpu():
    la, ga, _ = dc.run_train_step("cpu", False, B=2, H=128, W=128)
    lb, gb, _ = dc.run_train_step("cpu", False, B=2, H=128, W=128)
    assert set(la) == {"loss_rpn_cls", "loss_box_cls", "loss_box_reg", "loss_mask"}
    assert la["loss_mask"] > 0 and la["loss_box_reg"] > 0, "synthetic GT must give foreground proposals (all loss branches)"
    for k in la:
        assert abs(la[k] - lb[k]) <= 1e-6 * max(1.0, abs(la[k]))
    assert any(n.startswith("fpn.") for n in ga) and any(n.startswith("layer1.") for n in ga)
    assert max(dc.rel_err(ga[n], gb[n]) for n in ga) < 1e-5


def test_scripts_import_unchanged_with_stubs(tmp_path):
    metrics, val, info = dc.run_train_epoch("cpu", False, n_batches=1, B=2, H=128, W=128)
    assert np.isfinite(metrics["total_loss"]) and metrics["gradient_norm_mean"] > 0
    assert "mean_iou" in val and info["roi_align_type"].startswith("torchvision")
    status, shape, _ = dc.run_gradio_predict("cpu", False, tmp_path, H=96, W=128)
    assert status.startswith("Detected ") and len(shape) == 3
    preds, summary, _ = dc.run_tiles("cpu", False, tmp_path, n_tiles=1)
    assert preds[0]["masks"].dtype == np.uint8 and preds[0]["masks"].shape[1:] == (222, 300)


def test_rpn_sampling_host_logic_draws_the_reference_sample_on_cpu(monkeypatch):
    """matching.sample_rpn_anchors / rpn_compute_loss (the host logic above lcr_match_boxes_f32) against the reference's own
    RPN.compute_loss (src/components/rpn.py:42-123) from the same generator state.  There is no GPU here, so the KERNEL is
    stood in for by the oracle's match_boxes (test infrastructure; the product raises on CPU tensors —
    tests/test_gpu_dropin.py::test_fused_rpn_matching_draws_the_reference_sample is the real comparison)."""
    import importlib
    import torch
    from livecell_instance_segmentation_b200 import matching, ops
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
    from oracle import oracle

    def stand_in(a, g, pos_thr, neg_thr=None):
        mx, am, pos, neg, cnt = oracle.match_boxes(a.numpy(), g.numpy(), pos_thr, neg_thr)
        return tuple(torch.from_numpy(np.ascontiguousarray(v)) for v in (mx, am, pos, neg, cnt))
    with pytest.raises(Exception):
        ops.match_boxes(torch.zeros((1, 4)), torch.zeros((1, 4)), 0.5)          # the product has no CPU path
    monkeypatch.setattr(ops, "match_boxes", stand_in)
    ref_harness.import_reference()
    try:
        rpn_mod = importlib.import_module("src.components.rpn")
        anchors = importlib.import_module("src.components.anchor_generator").AnchorGenerator().generate_anchors((64, 64), stride=4, device="cpu")
        rpn = rpn_mod.RPN(in_channels=8, num_anchors=9)
        targets = dc.synth_targets(2, 256, 256, 5, "cpu")
        scores = torch.randn((2, 9, 64, 64), generator=torch.Generator().manual_seed(3))

        def run(fn, tg):
            s = scores.clone().requires_grad_(True)
            torch.manual_seed(77)
            loss = fn(rpn, [s], [None], anchors, tg, "cpu")["loss_rpn_cls"]
            loss.backward()
            return loss.detach(), s.grad
        (l_ref, g_ref), (l_got, g_got) = run(rpn_mod.RPN.compute_loss, targets), run(matching.rpn_compute_loss, targets)
        assert torch.equal(l_ref, l_got) and torch.equal(g_ref, g_got)
        assert int((g_ref != 0).sum()) == 256
        empty = [{"boxes": torch.zeros((0, 4))} for _ in range(2)]
        assert torch.equal(run(rpn_mod.RPN.compute_loss, empty)[0], run(matching.rpn_compute_loss, empty)[0])
    finally:
        ref_harness.purge()


def test_install_fused_rpn_matching_patches_and_restores_the_method():
    """install(fused_rpn_matching=True) rebinds RPN.compute_loss of the staged reference and uninstall() gives it back; the
    default install() leaves the method alone (patching only: nothing is launched on the CPU)."""
    import importlib
    from livecell_instance_segmentation_b200 import install as inst, matching
    ref_harness.import_reference()
    try:
        rpn_mod = importlib.import_module("src.components.rpn")
        original = rpn_mod.RPN.compute_loss
        done = inst.install()
        assert rpn_mod.RPN.compute_loss is original and "RPN.compute_loss" not in done.get("src.components.rpn", [])
        inst.uninstall()
        done = inst.install(fused_rpn_matching=True)
        assert "RPN.compute_loss" in done["src.components.rpn"] and rpn_mod.RPN.compute_loss is matching.rpn_compute_loss
        assert inst.uninstall() > 0 and rpn_mod.RPN.compute_loss is original
    finally:
        inst.uninstall()
        ref_harness.purge()

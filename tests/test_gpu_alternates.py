"""The alternative kernels kept behind tuning switches (A/B baselines and fallbacks for shapes the fast paths do not take)
run through the same parity checks as the defaults, inside the driver's `-m gpu` run (they were only covered by
tools/gpu_check.sh).  Switches go through lcr_set_tuning (the library reads the environment once)."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("roi_cpu_coords")]   # oracle / golden = torchvision's CPU-op coordinate rule


@pytest.fixture(scope="module")
def ops():
    from livecell_instance_segmentation_b200 import ops as o
    return o


def test_select_general_cluster_kernel(ops, oracle, synth, tune):
    """LCR_SELECT=general: the 8-CTA-cluster radix select (the path taken for > 8192 threshold survivors / NaNs) instead of
    the threshold-first prefilter + sort, on a crowded C3-sized map and a small ragged batch."""
    from gpu_util import N, T
    base = oracle.base_anchors()
    tune(LCR_SELECT="general")
    for (B, h, w, cells, k, H, W) in ((1, 130, 176, 2000, 2000, 520, 704), (3, 24, 32, 60, 200, 96, 128)):
        obj = synth.make_objectness(B, 9, h, w, n_cells=cells, seed=5 + h, k=k)
        gb, gs, gi, gc = ops.rpn_select([T(obj)], k=k, img_size=(H, W), score_thresh=0.3, min_size=10.0, strides=[4], base=torch.from_numpy(base))
        for b in range(B):
            ob, os_, oi = oracle.rpn_select(obj[b], base=base, k=k, score_thresh=0.3, min_size=10, img_h=H, img_w=W)
            n = int(gc[b, 0])
            assert n == len(oi) and np.array_equal(N(gi)[b, 0, :n], oi) and np.array_equal(N(gb)[b, 0, :n], ob)
            np.testing.assert_allclose(N(gs)[b, 0, :n], os_, rtol=0, atol=1e-6)


@pytest.mark.parametrize("resolve", ["serial"])
def test_nms_serial_resolve(ops, oracle, synth, tune, resolve):
    """LCR_NMS_RESOLVE=serial: the chunk-walk resolve (what segments above 2048 boxes always use) on small segments too."""
    from gpu_util import N, T
    tune(LCR_NMS_RESOLVE=resolve)
    rng = np.random.RandomState(8)
    for n in (33, 700, 2000):
        boxes = synth.make_rois(n, 60 + n)[:, 1:]
        scores = rng.rand(n).astype(np.float32)
        keep, kc = ops.nms_batched(T(boxes[None]), T(scores[None]), 0.5, post_n=n, cpu_threshold=True)
        assert np.array_equal(N(keep)[0, : int(kc[0])], oracle.nms(boxes, scores, 0.5))


@pytest.mark.parametrize("P", [7, 14])
def test_roi_align_per_cta_kernels(ops, oracle, synth, tune, P):
    """LCR_ROI_FWD=cta / LCR_ROI_BWD=cta: the per-CTA NHWC kernels (r01 design; still the path for maps narrower than 4 columns)."""
    from gpu_util import N, T, nhwc, assert_close_rel
    tune(LCR_ROI_FWD="cta", LCR_ROI_BWD="cta")
    feat = synth.make_features(2, 64, 30, 36, seed=9)
    rois = synth.make_rois(50, 10, img_h=120, img_w=144, batch=2, edge_cases=True)
    out = N(ops.roi_align_fwd([nhwc(T(feat))], [0.25], T(rois), None, (P, P), 2, False))
    assert_close_rel(out, oracle.roi_align_fwd(feat, rois, P, P, 0.25, 2, False), 1e-5)
    gout = np.random.RandomState(3).standard_normal(out.shape).astype(np.float32)
    gin = torch.empty((2, 64, 30, 36), device="cuda:0").contiguous(memory_format=torch.channels_last)
    ops.roi_align_bwd(T(gout), [gin], [0.25], T(rois), None, 2, False, zero_grad=True)
    assert_close_rel(N(gin), oracle.roi_align_bwd(gout, rois, feat.shape, 0.25, 2, False), 1e-5)


@pytest.mark.parametrize("mode", ["rows16", "single", "zeros_last"])
def test_paste_alternative_kernels(ops, oracle, synth, tune, mode):
    """LCR_PASTE=rows16 (one 16-byte store per thread: frame widths the bulk kernels do not take), single / zeros_last (the
    single-role bulk kernel of r01d and its r01c store order): bit-exact frames."""
    from gpu_util import N, T
    tune(LCR_PASTE=mode)
    for H, W, n in ((520, 704, 60), (96, 128, 25)):
        probs = synth.make_mask_probs(n, 28, 31 + W)
        boxes = synth.make_det_boxes(n, 32 + H, edge_cases=True) if H == 520 else synth.make_rois(n, 5, img_h=H, img_w=W, edge_cases=True)[:, 1:]
        valid = np.ones(n, np.uint8)
        valid[3] = 0
        buf = torch.full((n, H, W), 9, dtype=torch.uint8, device="cuda:0")
        out = N(ops.paste_masks(T(probs), T(boxes), H, W, valid=T(valid), out=buf))
        ref = oracle.paste_masks(probs, boxes, H, W)
        keep = valid.astype(bool)
        assert np.array_equal(out[keep], ref[keep]) and (out[~keep] == 9).all()


def test_ctypes_call_path_and_python_autograd_node(synth, oracle):
    """The call path used when csrc/lcr_torch.so is disabled (LCR_TORCH_EXT=0) or cannot be built: ctypes + the Python
    autograd.Function — same kernels, same results as the C++ autograd node."""
    from gpu_util import N, T, nhwc, assert_close_rel
    from livecell_instance_segmentation_b200 import ops, roi_align as ra
    feat = synth.make_features(1, 64, 24, 32, seed=12)
    rois = synth.make_rois(40, 13, img_h=96, img_w=128, edge_cases=True)
    gout = np.random.RandomState(4).standard_normal((40, 64, 7, 7)).astype(np.float32)
    res = []
    for f0 in (T(feat), nhwc(T(feat))):
        f = f0.clone().requires_grad_(True)
        y = ra._RoIAlignFn.apply(((0.25,), T(rois), None, (7, 7), 2, False), f)
        y.backward(T(gout))
        assert_close_rel(N(y), oracle.roi_align_fwd(feat, rois, 7, 7, 0.25, 2, False), 1e-5)
        assert_close_rel(N(f.grad), oracle.roi_align_bwd(gout, rois, feat.shape, 0.25, 2, False), 1e-5)
        res.append(N(y))
    f = T(feat).clone().requires_grad_(True)
    y = ra.roi_align(f, T(rois), (7, 7), 0.25, 2)                   # the default path (C++ node when the extension loads)
    assert_close_rel(N(y), res[0], 1e-6)
    boxes = rois[:, 1:]
    scores = np.random.RandomState(6).rand(40).astype(np.float32)
    keep, kc = ops.nms_batched(T(boxes[None]), T(scores[None]), ops._round_f32(0.4), post_n=40, cpu_threshold=True)
    assert np.array_equal(N(keep)[0, : int(kc[0])], N(ra.nms(T(boxes), T(scores), 0.4)))

"""GPU: no kernel writes outside the buffer it was given.  compute-sanitizer is closed on this pool (runs under it left
GPUs needing a reset), so every output of the hot path is placed inside a larger canary-filled allocation and the canaries
are checked afterwards — this is what caught nothing so far, and what would catch a mis-sized TMA bulk store."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("roi_cpu_coords")]   # oracle / golden = torchvision's CPU-op coordinate rule

PAD = 1 << 16      # bytes of canary on both sides


def guarded(nbytes, dtype, shape):
    raw = torch.full((PAD + nbytes + PAD,), 0xA5, dtype=torch.uint8, device="cuda:0")
    view = raw[PAD: PAD + nbytes].view(dtype).reshape(shape)
    return raw, view


def intact(raw, nbytes):
    return bool((raw[:PAD] == 0xA5).all()) and bool((raw[PAD + nbytes:] == 0xA5).all())


@pytest.mark.parametrize("P,K,C", [(7, 77, 256), (7, 5, 72), (14, 33, 256), (14, 9, 40)])
def test_roi_align_outputs_stay_in_bounds(synth, P, K, C):
    from gpu_util import T
    from livecell_instance_segmentation_b200 import ops
    feat = torch.randn((2, 40, 48, C), device="cuda:0").permute(0, 3, 1, 2)
    rois = T(synth.make_rois(K, 31, img_h=160, img_w=192, batch=2, edge_cases=True))
    nbytes = K * C * P * P * 4
    raw, out = guarded(nbytes, torch.float32, (K, C, P, P))
    ops.roi_align_fwd([feat], [0.25], rois, None, (P, P), 2, False, out=out)
    torch.cuda.synchronize()
    assert intact(raw, nbytes) and bool(torch.isfinite(out).all())
    gbytes = feat.numel() * 4
    graw, gin = guarded(gbytes, torch.float32, (2, 40, 48, C))
    gin = gin.permute(0, 3, 1, 2)
    ops.roi_align_bwd(torch.randn((K, C, P, P), device="cuda:0"), [gin], [0.25], rois, None, 2, False, zero_grad=True)
    torch.cuda.synchronize()
    assert intact(graw, gbytes) and bool(torch.isfinite(gin).all())


@pytest.mark.parametrize("H,W", [(520, 704), (222, 300), (64, 80), (33, 48)])
def test_paste_stays_in_bounds(synth, H, W):
    from gpu_util import T
    from livecell_instance_segmentation_b200 import ops
    n = 23
    boxes = T(synth.make_det_boxes(n, 41, img_h=H, img_w=W, lo=4, hi=min(H, W) - 2, edge_cases=True))
    probs = T(synth.make_mask_probs(n, 28, 42))
    nbytes = n * H * W
    raw, out = guarded(nbytes, torch.uint8, (n, H, W))
    ops.paste_masks(probs, boxes, H, W, out=out)
    torch.cuda.synchronize()
    assert intact(raw, nbytes)
    assert set(torch.unique(out).tolist()) <= {0, 255}


def test_select_nms_outputs_stay_in_bounds(synth):
    """rpn_select / nms / gather write only the capacities they were given (outputs are carved from one guarded arena)."""
    from gpu_util import T
    from livecell_instance_segmentation_b200 import ops
    obj = T(synth.make_objectness(3, 9, 24, 32, n_cells=80, seed=51, k=150))
    before = torch.cuda.memory_allocated()
    boxes, scores, index, counts = ops.rpn_select([obj], k=150, img_size=(96, 128), score_thresh=0.3, min_size=4.0, strides=[4],
                                                  base=ops.base_anchors())
    keep, kc = ops.nms_batched(boxes[:, 0], None, 0.4, post_n=60, counts=counts[:, 0].contiguous())
    torch.cuda.synchronize()
    assert int(counts.max()) <= 150 and int(kc.max()) <= 60
    c = counts[:, 0].tolist()
    for b in range(3):
        assert bool((index[b, 0, : c[b]] >= 0).all()) and bool((index[b, 0, : c[b]] < 9 * 24 * 32).all())
        k = int(kc[b])
        assert bool((keep[b, :k] >= 0).all()) and bool((keep[b, :k] < c[b]).all())
    assert torch.cuda.memory_allocated() >= before

"""GPU: size-independent properties at BASELINE.json's full sizes (where the CPU oracle would take minutes):
adjointness and linearity of RoIAlign fwd/bwd, NMS idempotence and the no-overlap invariant, top-k sortedness,
paste confinement — all through the C-ABI on the bench-sized tensors."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.usefixtures("roi_cpu_coords")]   # oracle / golden = torchvision's CPU-op coordinate rule


@pytest.fixture(scope="module")
def ops():
    from livecell_instance_segmentation_b200 import ops as o
    return o


@pytest.mark.parametrize("P,K,levels", [(7, 16384, 1), (7, 16384, 4), (14, 4096, 4)])
def test_roi_align_adjoint_and_linear(ops, synth, P, K, levels):
    """<roi_fwd(x), g> == <x, roi_bwd(g)> (backward is the exact transpose of forward) and forward is linear in x
    — C5 geometry, 256 channels, float64 accumulation of the dot products."""
    from gpu_util import T
    g = torch.Generator(device="cuda:0").manual_seed(5)
    shapes = [(130, 176), (65, 88), (33, 44), (17, 22)][:levels]
    scales = [0.25, 0.125, 0.0625, 0.03125][:levels]
    C = 256
    xs = [torch.randn((1, h, w, C), generator=g, device="cuda:0").permute(0, 3, 1, 2) for h, w in shapes]
    ys = [torch.randn((1, h, w, C), generator=g, device="cuda:0").permute(0, 3, 1, 2) for h, w in shapes]
    rois = T(synth.make_rois(K, 77, mode="anchor" if levels == 1 else "fpn", edge_cases=True))
    lvl = None if levels == 1 else ops.level_map(rois, 2, 5, 224.0, 4)
    fx = ops.roi_align_fwd(xs, scales, rois, lvl, (P, P), 2, False)
    gout = torch.randn(fx.shape, generator=g, device="cuda:0")
    grads = [torch.empty_like(x) for x in xs]
    ops.roi_align_bwd(gout, grads, scales, rois, lvl, 2, False, zero_grad=True)
    lhs = float((fx.double() * gout.double()).sum())
    rhs = float(sum((x.double() * gr.double()).sum() for x, gr in zip(xs, grads)))
    scale = float((fx.double().abs() * gout.double().abs()).sum())
    assert abs(lhs - rhs) <= 2e-6 * scale, (lhs, rhs, scale)
    fy = ops.roi_align_fwd(ys, scales, rois, lvl, (P, P), 2, False)
    comb = ops.roi_align_fwd([2.0 * x - 0.5 * y for x, y in zip(xs, ys)], scales, rois, lvl, (P, P), 2, False)
    ref = 2.0 * fx - 0.5 * fy
    assert float((comb - ref).abs().max()) <= 1e-5 * float(ref.abs().max())


def test_nms_invariants_at_bench_size(ops, synth):
    """64 segments x 2000 sorted candidates (the bench's proposal NMS): (1) no two kept boxes overlap above the
    threshold, (2) every dropped box overlaps an earlier kept box above it, (3) NMS of the kept set keeps everything."""
    import torchvision
    from gpu_util import T
    S, n, thr = 64, 2000, 0.4
    thr32 = float(np.float32(thr))   # the drop-in follows torchvision's CUDA op: IoU is compared with the threshold rounded to
    # fp32 (0.4f > 0.4).  Anchor-grid boxes DO produce pairs whose fp32 IoU equals 0.4f exactly; such a pair is kept.
    obj = T(np.concatenate([synth.make_objectness(1, 9, 130, 176, n_cells=2000, seed=300 + i, k=n) for i in range(4)]))
    obj = obj.repeat(S // 4, 1, 1, 1)
    boxes, scores, _, counts = ops.rpn_select([obj], k=n, img_size=(520, 704), score_thresh=0.3, min_size=10.0, strides=[4],
                                              base=ops.base_anchors())
    boxes, counts = boxes[:, 0].contiguous(), counts[:, 0].contiguous()
    keep, kc = ops.nms_batched(boxes, None, thr, post_n=n, counts=counts)
    for s in (0, 1, 2, 3, 63):
        c, k = int(counts[s]), int(kc[s])
        kept = keep[s, :k]
        assert bool((kept[1:] > kept[:-1]).all())                              # sorted input: kept indices ascending
        kb = boxes[s, kept]
        iou = torchvision.ops.box_iou(kb, kb).double()
        iou.fill_diagonal_(0)
        assert float(iou.max()) <= thr32                                        # (1)
        dropped = torch.ones(c, dtype=torch.bool, device="cuda:0")
        dropped[kept] = False
        di = torch.nonzero(dropped)[:, 0]
        cross = torchvision.ops.box_iou(boxes[s, di], kb).double()
        earlier = kept[None, :] < di[:, None]
        assert bool(((cross > thr32) & earlier).any(dim=1).all())               # (2)
        again, ka = ops.nms_batched(kb[None].contiguous(), None, thr, post_n=k)
        assert int(ka[0]) == k and bool((again[0, :k] == torch.arange(k, device="cuda:0")).all())   # (3)
    assert bool(torch.equal(kc[:4], kc[4:8])) and bool(torch.equal(kc[:4], kc[60:64]))             # identical segments agree
    for s in range(4):                                                                             # (entries past the count are unspecified)
        assert bool(torch.equal(keep[s, : int(kc[s])], keep[60 + s, : int(kc[s])]))


def test_select_outputs_sorted_and_thresholded(ops, synth):
    from gpu_util import T
    obj = T(np.concatenate([synth.make_objectness(1, 9, 130, 176, n_cells=2000, seed=400 + i, k=2000) for i in range(3)]))
    boxes, scores, index, counts = ops.rpn_select([obj], k=2000, img_size=(520, 704), score_thresh=0.3, min_size=10.0, strides=[4],
                                                  base=ops.base_anchors())
    for b in range(3):
        c = int(counts[b, 0])
        s = scores[b, 0, :c]
        assert c > 1500 and bool((s[:-1] >= s[1:]).all()) and float(s.min()) > 0.3
        assert int(torch.unique(index[b, 0, :c]).numel()) == c
        bx = boxes[b, 0, :c]
        assert float(bx.min()) >= 0 and float(bx[:, 2].max()) <= 704 and float(bx[:, 3].max()) <= 520
        assert bool(((bx[:, 2] - bx[:, 0]) >= 10).all()) and bool(((bx[:, 3] - bx[:, 1]) >= 10).all())
        # scores are the sigmoid of the logits the indices point at
        flat = obj[b].permute(1, 2, 0).reshape(-1)
        assert float((torch.sigmoid(flat[index[b, 0, :c]]) - s).abs().max()) <= 1e-6


def test_paste_confined_to_boxes_and_idempotent(ops, synth):
    from gpu_util import T
    n, H, W = 500, 520, 704
    boxes = T(synth.make_det_boxes(n, 8, edge_cases=True))
    probs = T(synth.make_mask_probs(n, 28, 9))
    a = ops.paste_masks(probs, boxes, H, W)
    b = ops.paste_masks(probs, boxes, H, W, out=torch.full((n, H, W), 77, dtype=torch.uint8, device="cuda:0"))
    assert torch.equal(a, b)                                                  # every byte is written, whatever was there
    ys = torch.arange(H, device="cuda:0")[None, :, None]
    xs = torch.arange(W, device="cuda:0")[None, None, :]
    bi = boxes.to(torch.int32)                                                # trunc, as box.int()
    inside = (xs >= bi[:, 0, None, None].clamp(min=0)) & (xs < bi[:, 2, None, None].clamp(max=W)) & \
             (ys >= bi[:, 1, None, None].clamp(min=0)) & (ys < bi[:, 3, None, None].clamp(max=H))
    assert not bool(((a > 0) & ~inside).any())
    assert set(torch.unique(a).tolist()) <= {0, 255}

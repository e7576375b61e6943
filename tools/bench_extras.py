"""Extra measurements bench.py carries beside `value`: every function times GPU work with
CUDA events on the current stream and returns plain dicts.  Nothing here touches oracle/ (CPU legs live in bench.py).

    roi_bwd_on_list      RoIAlign backward on the bench RoI list (zero-fill included)
    c5_sweep             BASELINE config C5: K x {1,4 levels} x {7,14} fwd + bwd, torchvision's sm_100 cubins beside it
    c2_train_shape       BASELINE config C2's RoIAlign: 128 RoIs on a 64x64 map, fwd + bwd through autograd
    config_latency       BASELINE configs C1 / C3: the whole region path of ONE frame as a CUDA graph
    match_boxes_c1       f1: anchors x ground truth row max + threshold masks + sums, torchvision's chain beside it
    h2d_ceiling          plain cudaMemcpyAsync of the e2e step's inputs from pinned memory (what PCIe gives this rank)
"""
import numpy as np


def _timed(fn, reps, flush=None, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        if flush is not None:
            flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def roi_bwd_on_list(ops, rois, n_live, F, C, FH, FW, peak, reps=5, seed=5):
    """lcr_roi_align_bwd_f32 on the bench step's RoI list: grad_out [K,C,7,7] -> grad maps [F,C,FH,FW] (channels_last),
    zero-fill of the maps inside the timed call.  Algorithmic bytes: grad_out read + 2 x the grad maps
    (read-modify-write of every touched line; crowded frames touch them all) + the zero-fill + 20 B per RoI."""
    import torch
    dev = rois.device
    K = rois.shape[0]
    g = torch.Generator(device=dev).manual_seed(seed)
    gout = torch.randn((K, C, 7, 7), generator=g, device=dev)
    gin = torch.empty((F, FH, FW, C), device=dev).permute(0, 3, 1, 2)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ms = _timed(lambda: ops.roi_align_bwd(gout, [gin], [0.25], rois, None, 2, False, zero_grad=True), reps, flush)
    ms_zero = _timed(lambda: gin.zero_(), reps, flush)
    nbytes = 4 * n_live * C * 49 + 3 * F * 4 * C * FH * FW + 20 * K
    # conservation: all-ones grad_out puts K_live * C * 49 into the maps (in-range RoIs)
    ones = torch.ones((K, C, 7, 7), device=dev)
    ops.roi_align_bwd(ones, [gin], [0.25], rois, None, 2, False, zero_grad=True)
    total = float(gin.double().sum())
    del gout, ones
    return {"ms": ms, "of_which_zero_fill_ms": ms_zero, "algorithmic_GB": nbytes / 1e9, "GBps": nbytes / 1e9 / (ms * 1e-3),
            "frac_of_hbm_peak": nbytes / 1e9 / (ms * 1e-3) / peak, "rois": int(n_live),
            "sum_check": {"sum_grad_in": total, "expected": float(n_live) * C * 49, "rel_err": abs(total - n_live * C * 49.0) / (n_live * C * 49.0)}}


def c5_sweep(ops, synth, peak, ks=(4096, 16384), reps=5, with_torchvision=True):
    """RoIs on the 704x520 pyramid (130x176 ... 17x22), 256 channels, 7x7 and 14x14, fwd + bwd."""
    import torch
    import torchvision
    dev = torch.device("cuda", torch.cuda.current_device())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    C = 256
    shapes = [(130, 176), (65, 88), (33, 44), (17, 22)]
    scales = [0.25, 0.125, 0.0625, 0.03125]
    g = torch.Generator(device=dev).manual_seed(7)
    feats_nhwc = [torch.randn((1, h, w, C), generator=g, device=dev).permute(0, 3, 1, 2) for h, w in shapes]
    feats_nchw = [f.contiguous() for f in feats_nhwc]
    rows = []
    for K in ks:
        for levels in (1, 4):
            rois = torch.from_numpy(synth.make_rois(K, 100 + K, mode="anchor" if levels == 1 else "fpn")).to(dev)
            lvl = None if levels == 1 else ops.level_map(rois, 2, 5, 224.0, 4)
            fl, sc = feats_nhwc[:levels], scales[:levels]
            feat_bytes = sum(4 * C * h * w for h, w in shapes[:levels])
            for P in (7, 14):
                out_bytes = 4 * K * C * P * P
                fwd_bytes = out_bytes + feat_bytes + 20 * K
                bwd_bytes = out_bytes + 3 * feat_bytes + 20 * K
                out = torch.empty((K, C, P, P), device=dev)
                ms_f = _timed(lambda: ops.roi_align_fwd(fl, sc, rois, lvl, (P, P), 2, False, out=out), reps, flush)
                gout = torch.randn((K, C, P, P), generator=g, device=dev)
                grads = [torch.empty_like(f) for f in fl]
                ms_b = _timed(lambda: ops.roi_align_bwd(gout, grads, sc, rois, lvl, 2, False, zero_grad=True), max(3, reps // 2), flush)
                row = {"K": K, "levels": levels, "P": P, "fwd_ms": ms_f, "fwd_frac_of_hbm_peak": fwd_bytes / 1e9 / (ms_f * 1e-3) / peak,
                       "bwd_ms": ms_b, "bwd_frac_of_hbm_peak": bwd_bytes / 1e9 / (ms_b * 1e-3) / peak,
                       "fwd_algorithmic_GB": fwd_bytes / 1e9, "bwd_algorithmic_GB": bwd_bytes / 1e9}
                if with_torchvision:
                    if levels == 1:
                        def tv():
                            return torchvision.ops.roi_align(feats_nchw[0], rois, (P, P), 0.25, 2, False)
                    else:
                        idx = [torch.where(lvl == l)[0] for l in range(levels)]
                        sub = [rois[i] for i in idx]

                        def tv():
                            res = torch.empty((K, C, P, P), device=dev)
                            for l in range(levels):
                                if len(idx[l]):
                                    res[idx[l]] = torchvision.ops.roi_align(feats_nchw[l], sub[l], (P, P), scales[l], 2, False)
                            return res
                    ms_tv = _timed(tv, max(3, reps // 2), flush)
                    ref = tv()
                    err = float((out - ref).abs().max() / ref.abs().max())
                    row.update(torchvision_cuda_fwd_ms=ms_tv, speedup_vs_torchvision=ms_tv / ms_f, max_rel_diff_vs_torchvision=err)
                    del ref
                rows.append(row)
                del out, gout, grads
    return rows


def c2_train_shape(synth, reps=20):
    """BASELINE config C2: image 0 of an 8 x 256x256 batch (64x64 level-0 map, NCHW as the FPN emits it), 128 sampled
    proposals, RoIAlign forward + backward through autograd (custom_maskrcnn.py:120, train_custom.py:44)."""
    import torch
    import torchvision
    from livecell_instance_segmentation_b200.roi_align import RoIAlign
    dev = torch.device("cuda", torch.cuda.current_device())
    g = torch.Generator(device=dev).manual_seed(9)
    C = 256
    feat = torch.randn((8, C, 64, 64), generator=g, device=dev).requires_grad_(True)
    boxes = torch.from_numpy(synth.make_rois(128, 5, img_h=256, img_w=256)[:, 1:]).to(dev)
    gout = torch.randn((128, C, 7, 7), generator=g, device=dev)
    op, optv = RoIAlign((7, 7), 0.25, 2), torchvision.ops.RoIAlign((7, 7), 0.25, 2)

    def step(o):
        def run():
            feat.grad = None
            o(feat[:1], [boxes]).backward(gout)
        return run
    ours = _timed(step(op), reps, warm=5)
    tv = _timed(step(optv), reps, warm=5)
    return {"fwd_plus_bwd_ms": ours, "torchvision_cuda_fwd_plus_bwd_ms": tv, "K": 128, "map": "1x256x64x64 NCHW",
            "note": "launch-latency-bound (6.4 MB out, 4.2 MB grad map): both are ~40 us of kernels inside the autograd engine"}


def match_boxes_c1(ops, synth, reps=20):
    """f1 at BASELINE config C1 size: 205 920 anchors x 160 ground-truth boxes.  Ours = lcr_match_boxes_f32 (one kernel:
    row max + `>= 0.5` / `< 0.3` masks + their sums); beside it the reference's chain as it runs on this GPU
    (torchvision.ops.box_iou, .max(dim=1), two compares, two sums — src/components/rpn.py:72-81), results compared."""
    import torch
    import torchvision
    dev = torch.device("cuda", torch.cuda.current_device())
    anc = ops.anchors(130, 176, 4, ops.base_anchors(), dev)
    gt = torch.from_numpy(synth.make_det_boxes(160, 3)).to(dev)

    def tv_chain():
        mx, _ = torchvision.ops.box_iou(anc, gt).max(dim=1)
        pos, neg = mx >= 0.5, mx < 0.3
        return pos, neg, pos.sum(), neg.sum()
    ours = _timed(lambda: ops.match_boxes(anc, gt, 0.5, 0.3), reps, warm=5)
    tv = _timed(tv_chain, reps, warm=5)
    _, _, pos, neg, cnt = ops.match_boxes(anc, gt, 0.5, 0.3)
    t_pos, t_neg, t_np, t_nn = tv_chain()
    same = bool(torch.equal(pos, t_pos) and torch.equal(neg, t_neg) and cnt.tolist() == [int(t_np.item()), int(t_nn.item())])
    return {"us": ours * 1e3, "torchvision_chain_us": tv * 1e3, "anchors": int(anc.shape[0]), "gt": 160,
            "positives": int(cnt[0].item()), "negatives": int(cnt[1].item()), "equal_to_torchvision_chain": same}


def config_latency(ops, synth, RegionConfig, RegionPipeline, reps=30):
    """BASELINE configs C1 (reference defaults: top-k 250 -> 50 proposals, ~150 cells) and C3 (crowded: 2000 cells, top-k
    2000 -> 1000 proposals -> 500 detections): the whole region path of ONE 704x520 frame, captured as one CUDA graph."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    H, W, h, w, C = 520, 704, 130, 176, 256
    out = {}
    for tag, cells, k, post, dets in (("c1", 150, 250, 50, 50), ("c3", 2000, 2000, 1000, 500)):
        pipe = RegionPipeline(RegionConfig(pre_nms_top_n=k, post_nms_top_n=post, max_detections=dets))
        obj = torch.from_numpy(synth.make_objectness(1, 9, h, w, n_cells=cells, seed=21, k=k)).to(dev)
        feat = torch.from_numpy(synth.make_features(1, C, h, w, seed=22)).to(dev).contiguous(memory_format=torch.channels_last)
        bs = torch.from_numpy(synth.make_box_scores((1, post), 23)).to(dev)
        probs = torch.from_numpy(synth.make_mask_probs(dets, 28, 24)).to(dev)
        masks = torch.empty((dets, H, W), dtype=torch.uint8, device=dev)
        roi_out = torch.empty((post, C, 7, 7), device=dev)

        def region():
            props = pipe.proposals(obj, (H, W))
            pipe.pool(feat, props.rois, out=roi_out)
            det = pipe.detections(props, bs)
            return props, pipe.paste(det, probs, (H, W), out=masks)

        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            for _ in range(3):
                region()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            props, det = region()
        ms = _timed(graph.replay, reps, warm=5)
        graph.replay()
        torch.cuda.synchronize()
        out[tag + "_latency_ms"] = ms
        out[tag] = {"proposals": int(props.counts.sum()), "detections": int(det.counts.sum()), "cells": cells,
                    "what": "rpn_select + NMS + gather + RoIAlign 7x7 + detection NMS + paste + records of one frame, one CUDA graph"}
        del graph
    return out


def h2d_ceiling(host, reps=3, streams=1):
    """What the host->device path gives THIS rank for the e2e step's inputs: plain cudaMemcpyAsync (Tensor.copy_ from
    pinned memory, one call per tensor, no batching API) into preallocated device buffers, nothing else running on the GPU."""
    import torch
    dev = torch.device("cuda", torch.cuda.current_device())
    dst = {k: torch.empty(v.shape, dtype=v.dtype, device=dev) for k, v in host.items()}
    nbytes = sum(int(v.numel() * v.element_size()) for v in host.values())
    ss = [torch.cuda.Stream() for _ in range(streams)]

    def copy_all():
        main = torch.cuda.current_stream()
        for i, k in enumerate(host):
            s = ss[i % streams]
            s.wait_stream(main)
            with torch.cuda.stream(s):
                dst[k].copy_(host[k], non_blocking=True)
        for s in ss:
            main.wait_stream(s)
    ms = _timed(copy_all, reps, warm=1)
    return {"GBps": nbytes / 1e9 / (ms * 1e-3), "ms": ms, "bytes": nbytes, "streams": streams}

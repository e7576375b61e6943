#!/usr/bin/env python
"""BASELINE configs C1 and C3: the region path on ONE 704x520 frame (latency view; bench.py is the throughput view).

    python tests/perf_config_latency.py > gpurun_out/config_latency.jsonl

(Lives under tests/ — not collected by pytest — because it times the CPU oracle next to the GPU path, and only
tests/, smoke() and bench.py's CPU legs may touch oracle/.)

C1: reference defaults (top-k 250 -> <= 50 proposals -> <= 50 detections, ~150 cells).
C3: crowded (2000 cells, top-k 2000 -> 1000 proposals -> 500 detections).
Per config: the batched pipeline captured in a CUDA graph (device time of the whole region path), the same issued
eagerly from Python, the per-image reference-interface shims (generate_inference_proposals -> RoIAlign -> nms ->
paste_masks_in_image, with the host syncs the reference API forces), the CPU oracle on the host cores, and — as
"the kernels to beat" — torchvision/ATen CUDA library calls for the three operators the reference delegates
(torch.topk over sigmoid scores, torchvision.ops.nms, torchvision.ops.roi_align) on the same tensors.
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def gpu_ms(fn, reps=30, warm=5):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def wall_ms(fn, reps=20, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e3


def main():
    import torch
    import torchvision
    from livecell_instance_segmentation_b200 import ops, synth
    from livecell_instance_segmentation_b200.pipeline import RegionConfig, RegionPipeline
    from livecell_instance_segmentation_b200.roi_align import RoIAlign, nms
    from livecell_instance_segmentation_b200.src.components.anchor_generator import AnchorGenerator
    from livecell_instance_segmentation_b200.src.utils.mask_utils import paste_masks_in_image
    from livecell_instance_segmentation_b200.src.utils.proposal_utils import generate_inference_proposals
    from oracle import oracle as orc
    orc.build()
    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    H, W, h, w, C = 520, 704, 130, 176, 256
    base = orc.base_anchors()
    for tag, cells, k, post, dets in (("C1", 150, 250, 50, 50), ("C3", 2000, 2000, 1000, 500)):
        cfg = RegionConfig(pre_nms_top_n=k, post_nms_top_n=post, max_detections=dets)
        pipe = RegionPipeline(cfg)
        obj_h = synth.make_objectness(1, 9, h, w, n_cells=cells, seed=21, k=k)
        feat_h = synth.make_features(1, C, h, w, seed=22)
        bs_h = synth.make_box_scores((1, post), 23)
        probs_h = synth.make_mask_probs(dets, 28, 24)
        obj, bs, probs = torch.from_numpy(obj_h).to(dev), torch.from_numpy(bs_h).to(dev), torch.from_numpy(probs_h).to(dev)
        feat = torch.from_numpy(feat_h).to(dev).contiguous(memory_format=torch.channels_last)
        masks = torch.empty((dets, H, W), dtype=torch.uint8, device=dev)
        roi_out = torch.empty((post, C, 7, 7), device=dev)

        def region():
            props = pipe.proposals(obj, (H, W))
            pipe.pool(feat, props.rois, out=roi_out)
            det = pipe.detections(props, bs)
            return pipe.paste(det, probs, (H, W), out=masks)

        eager = gpu_ms(region)
        side = torch.cuda.Stream()
        side.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(side):
            region()
        torch.cuda.current_stream().wait_stream(side)
        torch.cuda.synchronize()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            det = region()
        graphed = gpu_ms(graph.replay)
        graph.replay()
        torch.cuda.synchronize()
        n_det = int(det.counts.sum())
        n_prop = int(pipe.proposals(obj, (H, W)).counts.sum())

        # reference-interface shims, per image, dense outputs (host syncs included): wall clock
        gen = AnchorGenerator()
        roi_align = RoIAlign((7, 7), 0.25, 2)

        def shims():
            anc = gen.generate_anchors((h, w), 4, dev)
            p, s = generate_inference_proposals(obj[0], anc, (H, W), dev, num_pre_nms=k, num_post_nms=post)
            rf = roi_align(feat, [p])
            sc = bs[0, : len(p)]
            keep = sc > 0.4
            fb, fs = p[keep], sc[keep]
            kk = nms(fb, fs, 0.5)[:dets]
            return paste_masks_in_image(probs[: len(kk)], fb[kk], (H, W)), rf

        shim_ms = wall_ms(shims)

        # torchvision / ATen CUDA library calls on the same tensors (the operators the reference delegates)
        scores_flat = torch.sigmoid(obj[0]).permute(1, 2, 0).reshape(-1)
        tk = gpu_ms(lambda: torch.topk(torch.sigmoid(obj[0]).permute(1, 2, 0).reshape(-1), k))
        cand_boxes, cand_scores, _, cc = ops.rpn_select([obj], k=k, img_size=(H, W), score_thresh=0.3, min_size=10.0, strides=[4],
                                                        base=pipe.base)
        nb, ns = cand_boxes[0, 0, : int(cc[0, 0])].contiguous(), cand_scores[0, 0, : int(cc[0, 0])].contiguous()
        tv_nms = gpu_ms(lambda: torchvision.ops.nms(nb, ns, 0.4))
        our_nms_graph = torch.cuda.CUDAGraph()
        nbb, ncc = nb[None].contiguous(), cc[:, 0].contiguous()
        ops.nms_batched(nbb, None, 0.4, post_n=post, counts=ncc)
        torch.cuda.synchronize()
        with torch.cuda.graph(our_nms_graph):
            ops.nms_batched(nbb, None, 0.4, post_n=post, counts=ncc)
        our_nms = gpu_ms(our_nms_graph.replay)
        props = pipe.proposals(obj, (H, W))
        rois = props.rois[: int(props.counts[0])].contiguous()
        feat_nchw = feat.contiguous()
        tv_roi = gpu_ms(lambda: torchvision.ops.roi_align(feat_nchw, rois, (7, 7), 0.25, 2, False))
        our_roi = gpu_ms(lambda: ops.roi_align_fwd([feat], [0.25], rois, None, (7, 7), 2, False, out=roi_out))

        # CPU oracle, same frame
        def cpu():
            b_, s_, _ = orc.rpn_select(obj_h[0], base=base, k=k, score_thresh=0.3, min_size=10, img_h=H, img_w=W)
            keep = orc.nms(b_, None, 0.4, post_n=post)
            pb = b_[keep]
            r = np.concatenate([np.zeros((len(pb), 1), np.float32), pb], axis=1)
            orc.roi_align_fwd(feat_h, r)
            k2 = orc.nms(pb, bs_h[0, : len(pb)], 0.5, score_thresh=0.4, use_score_thresh=True, post_n=dets)
            orc.paste_masks(probs_h[: len(k2)], pb[k2], H, W)
        cpu()
        t0 = time.perf_counter()
        for _ in range(5):
            cpu()
        cpu_ms = (time.perf_counter() - t0) / 5 * 1e3
        print(json.dumps({"config": tag, "cells": cells, "proposals": n_prop, "detections": n_det,
                          "region_path_graph_ms": graphed, "region_path_eager_ms": eager, "reference_interface_shims_wall_ms": shim_ms,
                          "cpu_oracle_ms": cpu_ms, "cpu_threads": orc.num_threads(),
                          "nms_us": {"ours_graph": our_nms * 1e3, "torchvision_cuda": tv_nms * 1e3, "boxes": int(cc[0, 0])},
                          "roi_align_fwd_us": {"ours": our_roi * 1e3, "torchvision_cuda": tv_roi * 1e3, "rois": int(rois.shape[0])},
                          "topk_us": {"torch_sigmoid_topk": tk * 1e3, "ours": "inside rpn_select (threshold-first)"}}), flush=True)


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""RoIAlign forward 7x7 on ONE NCHW-contiguous 256 x 130 x 176 map (what the reference's FPN hands its pooler), K = 16..4096:
roi_fwd_planes_kernel (library default for NCHW) vs the one-thread-per-output generic kernel (LCR_ROI_FWD=generic) vs
transpose-to-NHWC + the NHWC fast path (what roi_align.RoIAlign picks above _prefer_nhwc's threshold) vs torchvision's CUDA op.
CUDA events, L2 flushed between repetitions; results checked against the generic kernel (bit-identical / 1e-6).

    python tools/roi_nchw_exp.py > gpurun_out/roi_nchw_exp.jsonl
"""
import json, os, sys
import numpy as np, torch, torchvision
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from livecell_instance_segmentation_b200 import ops, synth, _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))
C = 256
for mode, H, W in (("cells", 130, 176), ("anchor", 130, 176), ("cells", 64, 64)):   # 64 x 64: level 0 of a 256 x 256 training tile
    feat = torch.from_numpy(synth.make_features(1, C, H, W, seed=5)).to(dev)
    for K in (16, 50, 128, 256, 512, 1024, 4096):
        if mode == "cells":   # cell-sized proposals (the reference's own population after NMS): 16-76 px
            rois = np.concatenate([np.zeros((K, 1), np.float32), synth.make_det_boxes(K, 70 + K, img_h=4 * H, img_w=4 * W)], axis=1)
        else:
            rois = synth.make_rois(K, 100 + K, mode="anchor")
        rois = torch.from_numpy(rois).to(dev)
        out = torch.empty((K, C, 7, 7), device=dev)
        res = {}
        _lib.set_tuning("LCR_ROI_FWD", "generic")
        res["generic"] = timed(lambda: ops.roi_align_fwd([feat], [0.25], rois, None, (7, 7), 2, False, out=out))
        ref = out.clone()
        _lib.set_tuning("LCR_ROI_FWD", None)
        res["planes"] = timed(lambda: ops.roi_align_fwd([feat], [0.25], rois, None, (7, 7), 2, False, out=out))
        same = bool(torch.equal(out, ref))
        def graphed(fn, n=20):   # kernel time without the Python launch path: n calls in one CUDA graph (L2-warm)
            s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                fn()
            torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                for _ in range(n): fn()
            g.replay(); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n
        for ctas in (2, 4, 8, 16):
            _lib.set_tuning("LCR_ROI_PLANES_CTAS", ctas)
            res[f"planes_graph_ctas{ctas}"] = graphed(lambda: ops.roi_align_fwd([feat], [0.25], rois, None, (7, 7), 2, False, out=out))
        _lib.set_tuning("LCR_ROI_PLANES_CTAS", None)
        _lib.set_tuning("LCR_ROI_FWD", "generic")
        res["generic_graph"] = graphed(lambda: ops.roi_align_fwd([feat], [0.25], rois, None, (7, 7), 2, False, out=out))
        _lib.set_tuning("LCR_ROI_FWD", None)
        fn_buf = torch.empty_strided(feat.shape, (H * W * C, 1, W * C, C), device=dev)
        res["transpose_plus_nhwc_graph"] = graphed(lambda: ops.roi_align_fwd([ops.to_nhwc(feat, out=fn_buf)], [0.25], rois, None, (7, 7), 2, False, out=out))
        res["torchvision_graph"] = graphed(lambda: torchvision.ops.roi_align(feat, rois, (7, 7), 0.25, 2, False))
        res["transpose_plus_nhwc"] = timed(lambda: ops.roi_align_fwd([ops.to_nhwc(feat)], [0.25], rois, None, (7, 7), 2, False, out=out))
        err = float((out - ref).abs().max() / ref.abs().max())
        res["torchvision"] = timed(lambda: torchvision.ops.roi_align(feat, rois, (7, 7), 0.25, 2, False))
        print(json.dumps(dict(rois=mode, map=[H, W], K=K, **{k: round(v * 1e3, 1) for k, v in res.items()}, unit="us",
                              planes_equals_generic=same, nhwc_max_rel_diff=err)), flush=True)

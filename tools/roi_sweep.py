#!/usr/bin/env python
"""BASELINE config C5 (RoIAlign sweep) and the C2-shaped RoIAlign fwd/bwd, on one B200.

    python tools/roi_sweep.py [--reps 10] > gpurun_out/roi_sweep.jsonl

C5: K in {256, 1024, 4096, 16384} RoIs x {single level-0 map, 4 FPN levels}, 256 channels, 7x7 (box) and
14x14 (mask) pooling, forward and backward, 704x520 frame geometry (maps 130x176, 65x88, 33x44, 17x22).
C2: batch of 8 synthetic 256x256 tiles -> 64x64 level-0 map, 128 sampled RoIs on image 0 (the reference's training
step pools image 0 only, custom_maskrcnn.py:120), forward + backward.
Each line also carries torchvision's own CUDA op (the sm_100 cubins shipped in torchvision/_C.so, NCHW input) on
the same RoIs: the kernel to beat.  Algorithmic bytes as in SURVEY.md §8(d).
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def timed(fn, reps, flush):
    import torch
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--ks", default="256,1024,4096,16384")
    args = ap.parse_args()
    import torch
    import torchvision
    from livecell_instance_segmentation_b200 import ops, synth

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    C, H, W = 256, 520, 704
    shapes = [(130, 176), (65, 88), (33, 44), (17, 22)]
    scales = [0.25, 0.125, 0.0625, 0.03125]
    g = torch.Generator(device=dev).manual_seed(7)
    feats_nhwc = [torch.randn((1, h, w, C), generator=g, device=dev).permute(0, 3, 1, 2) for h, w in shapes]
    feats_nchw = [f.contiguous() for f in feats_nhwc]

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    for K in [int(k) for k in args.ks.split(",")]:
        for levels in (1, 4):
            rois = torch.from_numpy(synth.make_rois(K, 100 + K, mode="anchor" if levels == 1 else "fpn")).to(dev)
            lvl = None if levels == 1 else ops.level_map(rois, 2, 5, 224.0, 4)
            fl, sc = feats_nhwc[:levels], scales[:levels]
            feat_bytes = sum(4 * C * h * w for h, w in shapes[:levels])
            for P in (7, 14):
                out_bytes = 4 * K * C * P * P
                fwd_bytes = out_bytes + feat_bytes + 20 * K
                bwd_bytes = out_bytes + 3 * feat_bytes + 20 * K          # grad_out read + zero-fill + RMW of grad maps
                out = torch.empty((K, C, P, P), device=dev)
                ms_f = timed(lambda: ops.roi_align_fwd(fl, sc, rois, lvl, (P, P), 2, False, out=out), args.reps, flush)
                gout = torch.randn((K, C, P, P), generator=g, device=dev)
                grads = [torch.empty_like(f) for f in fl]               # channels_last like the features
                ms_b = timed(lambda: ops.roi_align_bwd(gout, grads, sc, rois, lvl, 2, False, zero_grad=True), max(3, args.reps // 2), flush)
                # torchvision on the same RoIs (per level, as MultiScaleRoIAlign does), NCHW
                if levels == 1:
                    tv = lambda: torchvision.ops.roi_align(feats_nchw[0], rois, (P, P), 0.25, 2, False)
                else:
                    idx = [torch.where(lvl == l)[0] for l in range(levels)]
                    sub = [rois[i] for i in idx]

                    def tv():
                        res = torch.empty((K, C, P, P), device=dev)
                        for l in range(levels):
                            if len(idx[l]):
                                res[idx[l]] = torchvision.ops.roi_align(feats_nchw[l], sub[l], (P, P), scales[l], 2, False)
                        return res
                ms_tv = timed(tv, max(3, args.reps // 2), flush)
                err = float((ops.roi_align_fwd(fl, sc, rois, lvl, (P, P), 2, False) - tv()).abs().max())
                emit(config="C5", K=K, levels=levels, P=P, fwd_ms=ms_f, fwd_GBps=fwd_bytes / 1e9 / (ms_f * 1e-3),
                     fwd_frac_hbm=fwd_bytes / 1e9 / (ms_f * 1e-3) / peak, bwd_ms=ms_b, bwd_GBps=bwd_bytes / 1e9 / (ms_b * 1e-3),
                     bwd_frac_hbm=bwd_bytes / 1e9 / (ms_b * 1e-3) / peak, torchvision_fwd_ms=ms_tv, speedup_vs_torchvision=ms_tv / ms_f,
                     max_abs_diff_vs_torchvision=err)
                del out, gout, grads

    # C2: training-step RoIAlign (image 0 of an 8 x 256x256 batch, 128 sampled proposals), fwd + bwd through autograd
    from livecell_instance_segmentation_b200.roi_align import RoIAlign
    feat = torch.randn((8, 64, 64, C), generator=g, device=dev).permute(0, 3, 1, 2).requires_grad_(True)
    boxes = torch.from_numpy(synth.make_rois(128, 5, img_h=256, img_w=256)[:, 1:]).to(dev)
    op = RoIAlign((7, 7), 0.25, 2)
    gout = torch.randn((128, C, 7, 7), generator=g, device=dev)

    def c2():
        feat.grad = None
        y = op(feat[:1], [boxes])
        y.backward(gout)
    ms = timed(c2, args.reps, flush)
    ftv = feat.detach().contiguous().requires_grad_(True)
    optv = torchvision.ops.RoIAlign((7, 7), 0.25, 2)

    def c2tv():
        ftv.grad = None
        y = optv(ftv[:1], [boxes])
        y.backward(gout)
    ms_tv = timed(c2tv, args.reps, flush)
    emit(config="C2", K=128, fwd_plus_bwd_ms=ms, torchvision_fwd_plus_bwd_ms=ms_tv, note="latency-bound: 6.4 MB out, 4.2 MB grad map")


if __name__ == "__main__":
    main()

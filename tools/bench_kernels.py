#!/usr/bin/env python
"""Per-kernel A/B timings at the bench workload (64 crowded 704x520 frames), CUDA events, one process.

    python tools/bench_kernels.py [--frames 64] [--reps 10] [--only roi,paste,select,nms]

Variants are switched through the library's tuning environment variables (read per call):
LCR_ROI_FWD=cta|warp, LCR_ROI_STREAM_OUT=0|1, LCR_PASTE=rows16|bulk.  Every variant's output is compared with
the first variant's (bit-exact for paste, 1e-5 relative for RoIAlign) so a faster-but-wrong kernel shows up here.
Prints one JSON line per measurement.
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B  # noqa: E402


def timed(fn, reps, flush=None):
    import torch
    for _ in range(3):
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
    return float(np.median(ms)), float(np.min(ms))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=64)
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only", default="roi,paste,select,nms,bwd,layout")
    args = ap.parse_args()
    only = set(args.only.split(","))
    import torch
    from livecell_instance_segmentation_b200 import _lib, ops
    from livecell_instance_segmentation_b200.pipeline import RegionConfig, RegionPipeline

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    F = args.frames
    peak, _ = B.peaks()
    obj_h, bs_h = B.make_host_inputs(F, seed0=0)
    probs_d = torch.from_numpy(B.make_mask_probs(F * B.MAX_DET, 99)).to(dev)
    g = torch.Generator(device=dev).manual_seed(1234)
    feat = torch.randn((F, B.FH, B.FW, B.C), generator=g, device=dev).permute(0, 3, 1, 2)
    obj_d, bs_d = torch.from_numpy(obj_h).to(dev), torch.from_numpy(bs_h).to(dev)
    pipe = RegionPipeline(RegionConfig(pre_nms_top_n=B.PRE_NMS, post_nms_top_n=B.POST_NMS, max_detections=B.MAX_DET))
    props = pipe.proposals(obj_d, (B.IMG_H, B.IMG_W))
    det = pipe.detections(props, bs_d)
    n_props, n_det = int(props.counts.sum()), int(det.counts.sum())
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)   # > L2: cold-cache timing between reps

    def emit(**kw):
        print(json.dumps(kw), flush=True)

    def setenv(env):
        for k in ("LCR_ROI_TEAM_PASS2", "LCR_ROI_XW", "LCR_ROI_RING_KB", "LCR_ROI_STAGED_WARPS", "LCR_ROI_RPC", "LCR_ROI_FWD", "LCR_ROI_STREAM_OUT", "LCR_PASTE", "LCR_ROI_BWD", "LCR_ROI_IPW", "LCR_SELECT", "LCR_ROI_SPLIT", "LCR_PASTE_ZB_KB", "LCR_PASTE_CTAS", "LCR_ROI_SHARED_TABLES", "LCR_NMS_SPAN",
                  "LCR_NMS_RESOLVE"):
            _lib.set_tuning(k, None)       # the library reads the environment only once: switches go through lcr_set_tuning
        for k, v in env.items():
            _lib.set_tuning(k, v)

    if "roi" in only:
        bytes_roi = 4 * n_props * B.C * 49 + F * 4 * B.C * B.FH * B.FW + 20 * F * B.POST_NMS
        ref = None
        S = {"LCR_ROI_FWD": "staged"}
        for name, env in [("sample walk (warp), ipw2", {"LCR_ROI_FWD": "warp"}),
                          ("default", {}),
                          ("persistent teams (team)", {"LCR_ROI_FWD": "team"}),
                          ("persistent teams, first pass only", {"LCR_ROI_FWD": "team", "LCR_ROI_TEAM_PASS2": "0"}),
                          ("row program, pipelined rows (rmp)", {"LCR_ROI_FWD": "rmp"}),
                          ("rmp, ipw1", {"LCR_ROI_FWD": "rmp", "LCR_ROI_IPW": "1"}),
                          ("rmp, ipw3", {"LCR_ROI_FWD": "rmp", "LCR_ROI_IPW": "3"}),
                          ("rmp, ipw4", {"LCR_ROI_FWD": "rmp", "LCR_ROI_IPW": "4"}),
                          ("row-major, 2 rows per phase (rm)", {"LCR_ROI_FWD": "rm"}),
                          ("row-major, 1 row per phase (rm1)", {"LCR_ROI_FWD": "rm1"}),
                          ("row-major rm, ipw1", {"LCR_ROI_FWD": "rm", "LCR_ROI_IPW": "1"}),
                          ("row-major rm, ipw3", {"LCR_ROI_FWD": "rm", "LCR_ROI_IPW": "3"}),
                          ("row-major rm, ipw4", {"LCR_ROI_FWD": "rm", "LCR_ROI_IPW": "4"}),
                          ("staged: 8 pooling warps, ring 56 KB (2 CTAs/SM), rpc 8", S),
                          ("staged: 4 pooling warps, ring 56 KB", {**S, "LCR_ROI_STAGED_WARPS": "4"}),
                          ("staged: 4 warps, ring 150 KB (1 CTA/SM), rpc 32", {**S, "LCR_ROI_RING_KB": "150", "LCR_ROI_RPC": "32", "LCR_ROI_STAGED_WARPS": "4"}),
                          ("staged kernel, nothing staged (consumers gather)", {"LCR_ROI_FWD": "staged_direct"}),
                          ("sample walk (warp) again", {"LCR_ROI_FWD": "warp"})]:
            setenv(env)
            out = pipe.pool(feat, props.rois)
            torch.cuda.synchronize()
            if ref is None:
                ref = out
                err = 0.0
            else:
                err = float((out - ref).abs().max() / ref.abs().max())
            med, mn = timed(lambda: pipe.pool(feat, props.rois), args.reps, flush)
            emit(kernel="roi_align_fwd", variant=name, ms=med, ms_min=mn, GBps=bytes_roi / 1e9 / (med * 1e-3),
                 frac=bytes_roi / 1e9 / (med * 1e-3) / peak, rel_err_vs_first=err, rois=n_props)
            del out
        del ref
        setenv({})

    if "roi1" in only:   # the RoIAlign forward selected by the caller's environment, nothing else (ncu target)
        med, mn = timed(lambda: pipe.pool(feat, props.rois), args.reps, flush)
        emit(kernel="roi_align_fwd", variant=os.environ.get("LCR_ROI_FWD", "default"), ms=med, ms_min=mn, rois=n_props)

    if "roiexp" in only:
        # what bounds the forward kernel?  same launch, different RoI lists (time per item vs window size / locality)
        r0 = props.rois.clone()
        live = r0[:, 0] >= 0
        def run(name, rois):
            med, mn = timed(lambda: pipe.pool(feat, rois), args.reps, flush)
            w = (rois[:, 3] - rois[:, 1])[live] * 0.25
            h = (rois[:, 4] - rois[:, 2])[live] * 0.25
            emit(kernel="roi_align_fwd experiment", variant=name, ms=med, ms_min=mn, mean_window_px=float(((w + 2) * (h + 2)).mean()))
        run("real list", r0)
        same = r0.clone()
        same[:, 1:] = torch.tensor([100.0, 100.0, 164.0, 164.0], device=dev)
        run("every RoI the same 64x64 box (all L1/L2 hits)", same)
        for side in (32.0, 64.0, 128.0):
            fx = r0.clone()
            cx, cy = (r0[:, 1] + r0[:, 3]) / 2, (r0[:, 2] + r0[:, 4]) / 2
            cx = cx.clamp(side / 2, B.IMG_W - side / 2)
            cy = cy.clamp(side / 2, B.IMG_H - side / 2)
            fx[:, 1], fx[:, 2], fx[:, 3], fx[:, 4] = cx - side / 2, cy - side / 2, cx + side / 2, cy + side / 2
            run(f"real centres, every box {int(side)}x{int(side)}", fx)
        pad = r0.clone()
        pad[:, 0] = -1.0
        run("every RoI padding (zero tile + bulk store only: the store path alone)", pad)
        tiny = r0.clone()
        cx, cy = (r0[:, 1] + r0[:, 3]) / 2, (r0[:, 2] + r0[:, 4]) / 2
        tiny[:, 1], tiny[:, 2], tiny[:, 3], tiny[:, 4] = cx, cy, cx + 2.0, cy + 2.0
        run("real centres, every box 2x2 px (tables + row loop + store, 2x2-pixel windows)", tiny)
        # spatially sorted inside each frame (row-major 64-px cells): neighbours in the list overlap
        key = r0[:, 0] * 1e6 + torch.floor((r0[:, 2] + r0[:, 4]) / 128) * 1e3 + (r0[:, 1] + r0[:, 3]) / 2
        key = torch.where(live, key, torch.full_like(key, 1e12))
        run("real boxes, spatially sorted per frame", r0[torch.argsort(key)].contiguous())

    if "bwd" in only:
        gout = torch.randn((F * B.POST_NMS, B.C, 7, 7), generator=g, device=dev)
        gin = torch.empty((F, B.C, B.FH, B.FW), device=dev).contiguous(memory_format=torch.channels_last)
        bytes_bwd = 4 * n_props * B.C * 49 + 3 * F * 4 * B.C * B.FH * B.FW + 20 * F * B.POST_NMS
        ref = None
        for name, env in [("cta(r01)", {"LCR_ROI_BWD": "cta"}), ("warp", {})]:
            setenv(env)
            ops.roi_align_bwd(gout, [gin], [0.25], props.rois, None, 2, False, zero_grad=True)
            torch.cuda.synchronize()
            if ref is None:
                ref, err = gin.clone(), 0.0
            else:
                err = float((gin - ref).abs().max() / ref.abs().max())
            med, mn = timed(lambda: ops.roi_align_bwd(gout, [gin], [0.25], props.rois, None, 2, False, zero_grad=True),
                            max(3, args.reps // 2), flush)
            emit(kernel="roi_align_bwd(+zero fill)", variant=name, ms=med, ms_min=mn, GBps=bytes_bwd / 1e9 / (med * 1e-3),
                 frac=bytes_bwd / 1e9 / (med * 1e-3) / peak, rel_err_vs_first=err)
        setenv({})
        del ref
        del gout, gin

    if "bwd1" in only:   # ncu target: the default backward on the bench list and on BASELINE C5 (16384 RoIs on ONE 130x176 map)
        from livecell_instance_segmentation_b200 import synth
        gout = torch.randn((F * B.POST_NMS, B.C, 7, 7), generator=g, device=dev)
        gin = torch.empty((F, B.C, B.FH, B.FW), device=dev).contiguous(memory_format=torch.channels_last)
        med, mn = timed(lambda: ops.roi_align_bwd(gout, [gin], [0.25], props.rois, None, 2, False, zero_grad=False), args.reps, flush)
        emit(kernel="roi_align_bwd (no zero fill)", variant="bench list", ms=med, rois=n_props)
        rois5 = torch.from_numpy(synth.make_rois(16384, 100 + 16384, mode="anchor")).to(dev)
        med, mn = timed(lambda: ops.roi_align_bwd(gout[:16384], [gin[:1]], [0.25], rois5, None, 2, False, zero_grad=False), args.reps, flush)
        emit(kernel="roi_align_bwd (no zero fill)", variant="C5: 16384 RoIs, one map", ms=med, rois=16384)
        del gout, gin

    if "paste" in only:
        masks = torch.empty((F * B.MAX_DET, B.IMG_H, B.IMG_W), dtype=torch.uint8, device=dev)
        boxes_flat = det.boxes.reshape(-1, 4)
        bytes_paste = n_det * (B.IMG_H * B.IMG_W + B.M * B.M * 4 + 16)
        ref = None
        for name, env in [("rows16(r01)", {"LCR_PASTE": "rows16"}), ("bulk, zero rows last (r01c)", {"LCR_PASTE": "zeros_last"}), ("bulk, single role (r01d)", {"LCR_PASTE": "single"}), ("bulk, warp-specialised", {}),
                          ("bulk, warp-specialised, 2 CTAs/SM", {"LCR_PASTE_CTAS": "2"}), ("bulk,zb8", {"LCR_PASTE_ZB_KB": "8"}),
                          ("bulk,zb32", {"LCR_PASTE_ZB_KB": "32"}), ("bulk,zb64", {"LCR_PASTE_ZB_KB": "64"})]:
            setenv(env)
            masks.fill_(7)
            ops.paste_masks(probs_d, boxes_flat, B.IMG_H, B.IMG_W, 0.5, 255, valid=det.valid, out=masks)
            torch.cuda.synchronize()
            if ref is None:
                ref = masks[: 8 * B.MAX_DET].clone()
                same = True
            else:
                same = bool(torch.equal(ref, masks[: 8 * B.MAX_DET]))
            med, mn = timed(lambda: ops.paste_masks(probs_d, boxes_flat, B.IMG_H, B.IMG_W, 0.5, 255, valid=det.valid, out=masks),
                            args.reps)
            emit(kernel="paste", variant=name, ms=med, ms_min=mn, GBps=bytes_paste / 1e9 / (med * 1e-3),
                 frac=bytes_paste / 1e9 / (med * 1e-3) / peak, identical_to_first=same, detections=n_det)
        setenv({})
        dead = torch.zeros_like(boxes_flat)
        med, mn = timed(lambda: ops.paste_masks(probs_d, dead, B.IMG_H, B.IMG_W, 0.5, 255, valid=det.valid, out=masks), args.reps)
        emit(kernel="paste experiment", variant="every box empty (zero-buffer bulk stores only)", ms=med, ms_min=mn,
             GBps=bytes_paste / 1e9 / (med * 1e-3), frac=bytes_paste / 1e9 / (med * 1e-3) / peak)
        med, mn = timed(lambda: masks.zero_(), args.reps)
        emit(kernel="reference: torch zero_() of the same 11.7 GB (pure HBM write stream)", ms=med, ms_min=mn,
             GBps=masks.numel() / 1e9 / (med * 1e-3), frac=masks.numel() / 1e9 / (med * 1e-3) / peak)
        del masks

    if "overlap" in only:
        # RoIAlign (L1/latency-bound) and paste (HBM-write-bound) side by side on two streams vs back to back
        masks = torch.empty((F * B.MAX_DET, B.IMG_H, B.IMG_W), dtype=torch.uint8, device=dev)
        boxes_flat = det.boxes.reshape(-1, 4)
        roi_out = torch.empty((F * B.POST_NMS, B.C, 7, 7), device=dev)
        s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()

        def roi():
            ops.roi_align_fwd([feat], [0.25], props.rois, None, (7, 7), 2, False, out=roi_out)

        def paste():
            ops.paste_masks(probs_d, boxes_flat, B.IMG_H, B.IMG_W, 0.5, 255, valid=det.valid, out=masks)

        def seq():
            roi()
            paste()

        def par():
            cur = torch.cuda.current_stream()
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            with torch.cuda.stream(s2):
                paste()
            with torch.cuda.stream(s1):
                roi()
            cur.wait_stream(s1)
            cur.wait_stream(s2)

        med, mn = timed(seq, args.reps, flush)
        emit(kernel="roi_align_fwd + paste", variant="back to back, one stream", ms=med, ms_min=mn)
        for cap in ("5", "4", "3", "2", "1"):
            setenv({"LCR_PASTE_CTAS": cap})
            med_p, _ = timed(paste, args.reps)
            med, mn = timed(par, args.reps, flush)
            emit(kernel="roi_align_fwd || paste", variant=f"two streams, paste capped at {cap} CTAs/SM", ms=med, ms_min=mn, paste_alone_ms=med_p)
        setenv({})
        del masks, roi_out

    if "layout" in only:
        # cost of an NCHW backbone: one tiled transpose of the level-0 maps before RoIAlign (channels_last models skip it)
        nchw = feat.contiguous()
        bytes_t = 2 * nchw.numel() * 4
        med, mn = timed(lambda: ops.to_nhwc(nchw), args.reps, flush)
        emit(kernel="nchw_to_nhwc (64 x 256 x 130 x 176)", variant="tiled transpose", ms=med, ms_min=mn, GBps=bytes_t / 1e9 / (med * 1e-3),
             frac=bytes_t / 1e9 / (med * 1e-3) / peak)
        del nchw

    if "select" in only:
        bytes_sel = F * (4 * B.A * B.FH * B.FW) + F * B.PRE_NMS * 28
        fn = lambda: ops.rpn_select([obj_d], k=B.PRE_NMS, img_size=(B.IMG_H, B.IMG_W), score_thresh=0.3, min_size=10.0, strides=[4],
                                    base=pipe.base)
        ref = None
        for name, env in [("general(r01)", {"LCR_SELECT": "general"}), ("threshold-first", {})]:
            setenv(env)
            outs = fn()
            torch.cuda.synchronize()
            if ref is None:
                ref, same = outs, True
            else:
                cnt = ref[3].reshape(-1).tolist()
                same = bool(torch.equal(ref[3], outs[3])) and all(
                    torch.equal(ref[j].reshape(len(cnt), B.PRE_NMS, -1)[s, :c], outs[j].reshape(len(cnt), B.PRE_NMS, -1)[s, :c])
                    for j in range(3) for s, c in enumerate(cnt))
            med, mn = timed(fn, args.reps, flush)
            emit(kernel="rpn_select", variant=name, ms=med, ms_min=mn, GBps=bytes_sel / 1e9 / (med * 1e-3),
                 frac=bytes_sel / 1e9 / (med * 1e-3) / peak, identical_to_first=same)
        setenv({})

    if "nms" in only:
        boxes, scores, _, counts = ops.rpn_select([obj_d], k=B.PRE_NMS, img_size=(B.IMG_H, B.IMG_W), score_thresh=0.3, min_size=10.0,
                                                  strides=[4], base=pipe.base)
        bx, ct = boxes[:, 0], counts[:, 0].contiguous()
        for name, env in [("serial resolve", {"LCR_NMS_RESOLVE": "serial"}), ("parallel resolve", {})]:
            setenv(env)
            med, mn = timed(lambda: ops.nms_batched(bx, None, 0.4, post_n=B.POST_NMS, counts=ct), args.reps)
            emit(kernel="nms(64 segments x 2000)", variant=name, ms=med, ms_min=mn)
            med, mn = timed(lambda: ops.nms_batched(props.boxes, bs_d, 0.5, post_n=B.MAX_DET, counts=props.counts, score_thresh=0.4), args.reps)
            emit(kernel="det nms(64 segments x 1000, scores)", variant=name, ms=med, ms_min=mn)
        setenv({})
        med, mn = timed(lambda: ops.nms_batched(bx[:1], None, 0.4, post_n=B.POST_NMS, counts=ct[:1]), args.reps)
        emit(kernel="nms(1 segment x 2000)", variant="default", us=med * 1e3, us_min=mn * 1e3)
        # single-segment latency without the Python launch path: the three launches replayed from a CUDA graph
        bx1, ct1 = bx[:1].contiguous(), ct[:1].contiguous()
        for name, env in [("serial resolve, span8", {"LCR_NMS_RESOLVE": "serial", "LCR_NMS_SPAN": "8"}),
                          ("parallel resolve, span8", {"LCR_NMS_SPAN": "8"}), ("parallel resolve, span16", {"LCR_NMS_SPAN": "16"}),
                          ("parallel resolve, span32", {"LCR_NMS_SPAN": "32"})]:
            setenv(env)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                for _ in range(3):
                    ops.nms_batched(bx1, None, 0.4, post_n=B.POST_NMS, counts=ct1)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
            gr = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gr):
                keep_g, kc_g = ops.nms_batched(bx1, None, 0.4, post_n=B.POST_NMS, counts=ct1)
            med, mn = timed(gr.replay, max(args.reps, 50))
            emit(kernel="nms(1 segment x 2000), graph replay", variant=name, us=med * 1e3, us_min=mn * 1e3, kept=int(kc_g[0]))
        setenv({})
        med, mn = timed(lambda: ops.nms_batched(props.boxes, bs_d, 0.5, post_n=B.MAX_DET, counts=props.counts, score_thresh=0.4), args.reps)
        emit(kernel="det nms(64 segments x 1000, scores)", variant="default", ms=med, ms_min=mn)


if __name__ == "__main__":
    main()

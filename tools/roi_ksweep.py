#!/usr/bin/env python
"""RoIAlign forward 7x7 over K in 64..65536, one frame and eight frames, one level (anchor-shaped RoIs up to 181 px) and four
FPN levels (level-mapped RoIs): library default vs the sample-walk kernel (LCR_ROI_FWD=warp) vs the persistent two-pass
kernel (team).  CUDA events, L2 flushed between repetitions.

    python tools/roi_ksweep.py > gpurun_out/roi_ksweep.jsonl
"""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from livecell_instance_segmentation_b200 import ops, synth, _lib
dev = torch.device("cuda:0")
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def timed(fn, reps=8):
    for _ in range(3): fn()
    torch.cuda.synchronize(); ms = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ms.append(e0.elapsed_time(e1))
    return float(np.median(ms))
C = 256
shapes = [(130, 176), (65, 88), (33, 44), (17, 22)]
scales = [0.25, 0.125, 0.0625, 0.03125]
g = torch.Generator(device=dev).manual_seed(7)
for B in (1, 8):
    feats = [torch.randn((B, h, w, C), generator=g, device=dev).permute(0, 3, 1, 2) for h, w in shapes]
    for levels in (1, 4):
        for K in (64, 256, 1024, 2048, 4096, 16384, 65536):
            if B == 1 and K > 16384: continue
            rois = torch.from_numpy(synth.make_rois(K, 100 + K, mode="anchor" if levels == 1 else "fpn", batch=B)).to(dev)
            if B > 1:
                rois = rois[torch.argsort(rois[:, 0], stable=True)].contiguous()
            lvl = None if levels == 1 else ops.level_map(rois, 2, 5, 224.0, 4)
            out = torch.empty((K, C, 7, 7), device=dev)
            res = {}
            for name in (None, "warp", "team"):
                _lib.set_tuning("LCR_ROI_FWD", name)
                res[str(name)] = timed(lambda: ops.roi_align_fwd(feats[:levels], scales[:levels], rois, lvl, (7, 7), 2, False, out=out))
            print(json.dumps(dict(B=B, levels=levels, K=K, **{k: round(v * 1e3, 1) for k, v in res.items()}, unit="us")), flush=True)

#!/usr/bin/env python
"""Two launches of roi_fwd_planes_kernel on one NCHW 256 x 130 x 176 map (K = 50: config C1; K = 1024) for ncu
(tools/gpu_profile_planes.sh).  Prints the eager times; checks against the per-output generic kernel."""
import os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from livecell_instance_segmentation_b200 import ops, synth, _lib
dev = torch.device("cuda:0")
feat = torch.from_numpy(synth.make_features(1, 256, 130, 176, seed=5)).to(dev)
for K in (50, 1024):
    rois = torch.from_numpy(np.concatenate([np.zeros((K, 1), np.float32), synth.make_det_boxes(K, 70 + K)], axis=1)).to(dev)
    _lib.set_tuning("LCR_ROI_FWD", "generic")
    ref = ops.roi_align_fwd([feat], [0.25], rois, None, (7, 7), 2, False)
    _lib.set_tuning("LCR_ROI_FWD", None)
    for _ in range(3):
        out = ops.roi_align_fwd([feat], [0.25], rois, None, (7, 7), 2, False)
    torch.cuda.synchronize()
    print(K, "equal to the generic kernel:", bool(torch.equal(out, ref)))

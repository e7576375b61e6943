"""Configurations + seeded inputs of the torchvision-RPN fixtures (tests/golden/tv_rpn.npz).  Imported by make_golden.py (which
runs torchvision on them) and by the tests (which regenerate the same inputs): only torchvision's OUTPUTS are stored."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from livecell_instance_segmentation_b200 import synth  # noqa: E402


def tv_rpn_case(tag):
    """Configuration + seeded inputs of one torchvision-RPN fixture (shared with the tests: inputs are regenerated from the
    seeds, only torchvision's outputs are stored).  Image sizes are multiples of 32, as GeneralizedRCNNTransform pads them,
    so that stride_h == stride_w on every level."""
    cases = {
        # boxes.numel() <= 4000 on CPU -> _batched_nms_coordinate_trick (TV:ops/boxes.py:83-86): what torchvision runs on a GPU
        # for up to 25 000 boxes per image
        "trick": dict(H=256, W=320, A=3, sizes=((32,), (64,), (128,), (256,)), ratios=(0.5, 1.0, 2.0), k=200, post=300, nms=0.7,
                      score_thresh=0.0, min_size=1e-3, seed=500, cells=(60, 40, 20, 8)),
        # > 4000 -> _batched_nms_vanilla (per-level nms); a score threshold that bites; 5 levels like maskrcnn_resnet50_fpn
        "vanilla": dict(H=256, W=320, A=3, sizes=((32,), (64,), (128,), (256,), (512,)), ratios=(0.5, 1.0, 2.0), k=400, post=250, nms=0.7,
                        score_thresh=0.05, min_size=1e-3, seed=520, cells=(120, 60, 30, 10, 4)),
    }
    c = dict(cases[tag])
    L = len(c["sizes"])
    c["shapes"] = [(c["H"] >> (2 + l), c["W"] >> (2 + l)) for l in range(L)]
    c["strides"] = [4 << l for l in range(L)]
    B = 2
    c["B"] = B
    c["obj"] = [synth.make_objectness(B, c["A"], h, w, n_cells=c["cells"][l], seed=c["seed"] + l, k=min(c["k"], c["A"] * h * w))
                for l, (h, w) in enumerate(c["shapes"])]
    rng = np.random.RandomState(c["seed"] + 99)
    c["deltas"] = [rng.normal(0.0, 0.3, size=(B, 4 * c["A"], h, w)).astype(np.float32) for (h, w) in c["shapes"]]
    return c

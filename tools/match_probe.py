#!/usr/bin/env python
"""lcr_match_boxes_f32 at BASELINE config C1 size (205 920 anchors x 160 ground-truth boxes) and at the C2 training shape
(36 864 anchors x 480): device time per launch with the launch queue kept full (events around 50 back-to-back calls), the
reference's chain (torchvision.ops.box_iou -> max -> two compares -> two sums) beside it.  Also the ncu target of
tools/gpu_profile_match.sh."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    import torch
    import torchvision
    from livecell_instance_segmentation_b200 import ops, synth
    dev = torch.device("cuda:0")
    for name, (h, w, G) in {"C1_130x176_x160": (130, 176, 160), "C2_64x64_x480": (64, 64, 480)}.items():
        anc = ops.anchors(h, w, 4, ops.base_anchors(), dev)
        gt = torch.from_numpy(synth.make_det_boxes(G, 3)).to(dev)

        def tv_chain():
            mx, _ = torchvision.ops.box_iou(anc, gt).max(dim=1)
            pos, neg = mx >= 0.5, mx < 0.3
            return pos.sum(), neg.sum()

        def timed(fn, n=50):
            for _ in range(5):
                fn()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(n):
                fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1) / n * 1e3
        ours = timed(lambda: ops.match_boxes(anc, gt, 0.5, 0.3))
        tv = timed(tv_chain)
        pairs = anc.shape[0] * G
        print(json.dumps({"case": name, "anchors": int(anc.shape[0]), "gt": G, "match_boxes_us": ours, "torchvision_chain_us": tv,
                          "pairs_per_ns": pairs / (ours * 1e3), "iou_matrix_MB_not_written": pairs * 4 / 1e6}), flush=True)


if __name__ == "__main__":
    main()

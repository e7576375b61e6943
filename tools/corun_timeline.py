#!/usr/bin/env python
"""Timeline of the streamed step (bench workload, CUDA graphs): when does each sub-batch's front group (select / NMS / gather /
RoIAlign / detection NMS, main stream) and back group (paste + records, side stream) start and end in steady state, against
their durations alone?  CUDA events around every graph replay.

    python tools/corun_timeline.py [--chunks 2] [--depth 2] > gpurun_out/corun_timeline.json
"""
import argparse, json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench as B
from livecell_instance_segmentation_b200.pipeline import RegionConfig, StreamedRegionPipeline

ap = argparse.ArgumentParser()
ap.add_argument("--chunks", type=int, default=2)
ap.add_argument("--depth", type=int, default=2)
ap.add_argument("--frames", type=int, default=64)
ap.add_argument("--batches", type=int, default=12)
args = ap.parse_args()
dev = torch.device("cuda:0")
F = args.frames
cfg = RegionConfig(pre_nms_top_n=B.PRE_NMS, post_nms_top_n=B.POST_NMS, max_detections=B.MAX_DET)
obj_h, bs_h = B.make_host_inputs(F, seed0=0)
probs_h = B.make_mask_probs(F * B.MAX_DET, 99)
g = torch.Generator(device=dev).manual_seed(1234)
sp = StreamedRegionPipeline(cfg, F, (B.C, B.FH, B.FW), (B.IMG_H, B.IMG_W), num_anchors=B.A, chunks=args.chunks, device=dev, depth=args.depth)
sp.inputs["obj"].copy_(torch.from_numpy(obj_h))
sp.inputs["bs"].copy_(torch.from_numpy(bs_h))
sp.inputs["probs"].copy_(torch.from_numpy(probs_h))
sp.inputs["feat"].copy_(torch.randn((F, B.FH, B.FW, B.C), generator=g, device=dev).permute(0, 3, 1, 2))
sp.run(finish=True)
sp.capture()
torch.cuda.synchronize()
main = torch.cuda.current_stream(dev)
E = lambda: torch.cuda.Event(enable_timing=True)

# each group alone (same graphs, one stream busy at a time)
alone = {"front": [], "back": []}
for _ in range(3):
    for slot in range(sp.chunks):
        for name, idx in (("front", 0), ("back", 1)):
            e0, e1 = E(), E()
            e0.record(main); sp._graphs[slot][idx].replay(); e1.record(main)
            torch.cuda.synchronize()
            alone[name].append(e0.elapsed_time(e1))

# the streamed schedule of StreamedRegionPipeline.run, with events around every replay
t0 = E(); t0.record(main)
marks = []
for b in range(args.batches):
    base = (b % sp.depth) * sp.chunks
    for c in range(sp.chunks):
        slot = base + c
        if sp.used[slot]:
            main.wait_event(sp.ev_side[slot])
        f0, f1, b0, b1 = E(), E(), E(), E()
        f0.record(main); sp._graphs[slot][0].replay(); f1.record(main)
        sp.ev_main[slot].record(main)
        with torch.cuda.stream(sp.side):
            sp.side.wait_event(sp.ev_main[slot])
            b0.record(sp.side); sp._graphs[slot][1].replay(); b1.record(sp.side)
            sp.ev_side[slot].record(sp.side)
        sp.used[slot] = True
        marks.append((b, c, f0, f1, b0, b1))
main.wait_stream(sp.side)
t1 = E(); t1.record(main)
torch.cuda.synchronize()
rows = [dict(batch=b, sub=c, front=[round(t0.elapsed_time(f0), 3), round(t0.elapsed_time(f1), 3)],
             back=[round(t0.elapsed_time(b0), 3), round(t0.elapsed_time(b1), 3)],
             front_ms=round(f0.elapsed_time(f1), 3), back_ms=round(b0.elapsed_time(b1), 3)) for b, c, f0, f1, b0, b1 in marks]
steady = rows[len(rows) // 2:]
print(json.dumps({
    "chunks": sp.chunks, "depth": sp.depth, "frames": F,
    "alone_ms": {k: round(float(np.median(v)), 3) for k, v in alone.items()},
    "streamed_ms_per_batch": round(t0.elapsed_time(t1) / args.batches, 3),
    "steady_front_ms": round(float(np.mean([r["front_ms"] for r in steady])), 3),
    "steady_back_ms": round(float(np.mean([r["back_ms"] for r in steady])), 3),
    "timeline": rows[-3 * sp.chunks:],
}))

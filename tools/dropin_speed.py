#!/usr/bin/env python
"""What a user of the reference sees after `install()`: wall-clock time of the UNMODIFIED reference model's own entry points
(staged checkout baseline/_ref), untouched vs patched, on one B200.

    python tools/dropin_speed.py > gpurun_out/dropin_speed.jsonl

* forward_inference (src/custom_maskrcnn.py:144-209) on 1 and 8 synthetic 704x520 frames (BASELINE config C1: reference
  defaults, 250 -> 50 proposals per frame);
* forward_train + backward (src/custom_maskrcnn.py:85-142) on 8 x 256x256 tiles (BASELINE config C2).
The backbone, FPN and heads are the reference's own PyTorch modules in both arms; only the region path changes.  Wall clock
with a device synchronize on both sides (the reference's loops are full of host syncs, so device time alone would flatter it).
"""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def wall_ms(fn, reps, warm=3):
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
    import numpy as np
    import torch
    import dropin_cases as dc
    import ref_harness
    from livecell_instance_segmentation_b200 import _lib, install as inst
    if not ref_harness.available():
        print(json.dumps({"unavailable": "baseline/_ref not staged"}))
        return
    dev = "cuda:0"
    res = {}
    for patched in (False, True, "batched", "fused"):
        cm = ref_harness.import_reference()
        if patched:
            inst.install(batched_inference=(patched == "batched"), fused_rpn_matching=(patched == "fused"))
        try:
            tag = {False: "untouched", True: "patched", "batched": "patched_batched_inference",
                   "fused": "patched_fused_rpn_matching"}[patched]
            model = dc.new_model(cm, dev, 0).eval()
            imgs1 = torch.from_numpy(np.stack([dc.synth_image(520, 704, 150, 1)])).to(dev)
            dc.calibrate_rpn(model, imgs1)
            imgs8 = torch.from_numpy(np.stack([dc.synth_image(520, 704, 150, 10 + b) for b in range(8)])).to(dev)
            l0 = _lib.launch_count()
            with torch.no_grad():
                n_det = sum(len(p["boxes"]) for p in model(imgs8))
                res[f"{tag}_inference_1x704x520_ms"] = wall_ms(lambda: model(imgs1), 20)
                res[f"{tag}_inference_8x704x520_ms"] = wall_ms(lambda: model(imgs8), 10)
                feats, _ = model.extract_features(imgs8)
                res[f"{tag}_backbone_fpn_only_8x704x520_ms"] = wall_ms(lambda: model.extract_features(imgs8), 10)
            res[f"{tag}_detections_in_8_frames"] = n_det
            # training step (C2)
            tmodel = dc.new_model(cm, dev, 0).train()
            timgs = torch.from_numpy(np.stack([dc.synth_image(256, 256, 40, 7 + b) for b in range(8)])).to(dev)
            targets = dc.synth_targets(8, 256, 256, 5, dev)
            dc.calibrate_rpn(tmodel, timgs)
            opt = torch.optim.AdamW(tmodel.parameters(), lr=1e-4)

            def step():
                loss = sum(v for v in tmodel(timgs, targets).values())
                opt.zero_grad()
                loss.backward()
                opt.step()
            res[f"{tag}_train_step_8x256x256_ms"] = wall_ms(step, 10)
            res[f"{tag}_liblcr_launches"] = _lib.launch_count() - l0
        finally:
            if patched:
                inst.uninstall()
            ref_harness.purge()
    for k in ("inference_1x704x520_ms", "inference_8x704x520_ms", "train_step_8x256x256_ms"):
        res["speedup_" + k[:-3]] = res["untouched_" + k] / res["patched_" + k]
        res["speedup_batched_" + k[:-3]] = res["untouched_" + k] / res["patched_batched_inference_" + k]
    res["speedup_fused_rpn_matching_train_step_8x256x256"] = (res["untouched_train_step_8x256x256_ms"]
                                                              / res["patched_fused_rpn_matching_train_step_8x256x256_ms"])
    res["note"] = ("wall clock, device synchronised; same weights and inputs; the per-image Python loop of forward_inference, the backbone, FPN "
                   "and heads are the reference's own code in both arms")
    print(json.dumps(res), flush=True)


if __name__ == "__main__":
    main()

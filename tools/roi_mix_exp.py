#!/usr/bin/env python
"""RoIAlign forward on lists that MIX row-program RoIs (<= 28 feature px) with large ones (sample walk): one 130x176 map,
16 384 anchor-shaped RoIs (sizes 32/64/128 x ratios .5/1/2), in random order, sorted by kind, and each kind alone.

    python tools/roi_mix_exp.py > gpurun_out/roi_mix_exp.jsonl

Columns: None = library default, warp = sample-walk kernel, team = persistent two-pass kernel, rm = row program, two rows
in flight.  What it showed (profiles/r02b_roi_mix_exp.jsonl): kernels that inline many loop bodies lose on MIXED lists
(instruction-cache misses: ncu stall_no_inst 56 % for the first team kernel) and not on the same list sorted by kind.
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
g = torch.Generator(device=dev).manual_seed(7)
K = 16384
feat = torch.randn((1, 130, 176, C), generator=g, device=dev).permute(0, 3, 1, 2)
base = synth.make_rois(K, 100, mode="anchor", batch=1)
w = base[:, 3] - base[:, 1]; h = base[:, 4] - base[:, 2]
slow = (w > 112) | (h > 100)
variants = {"random": base, "fast first, then slow": np.concatenate([base[~slow], base[slow]]), "slow first": np.concatenate([base[slow], base[~slow]]),
            "alternating blocks of 592": None, "only fast part": base[~slow], "only slow part": base[slow]}
f, sl = base[~slow], base[slow]
blocks = []
i = j = 0
while i < len(f) or j < len(sl):
    blocks.append(f[i:i + 1184]); i += 1184
    blocks.append(sl[j:j + 592]); j += 592
variants["alternating blocks of 592"] = np.concatenate(blocks)
for name, r in variants.items():
    rois = torch.from_numpy(np.ascontiguousarray(r)).to(dev)
    out = torch.empty((len(r), C, 7, 7), device=dev)
    res = {}
    for v in (None, "warp", "team", "rm"):
        _lib.set_tuning("LCR_ROI_FWD", v)
        res[str(v)] = round(timed(lambda: ops.roi_align_fwd([feat], [0.25], rois, None, (7, 7), 2, False, out=out)) * 1e3, 1)
    print(json.dumps(dict(list=name, n=len(r), **res)), flush=True)

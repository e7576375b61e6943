#!/usr/bin/env python
"""bench.py — images/s through RPN select -> NMS -> RoIAlign -> detection NMS -> mask paste.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...      (N > 1)

Workload (BASELINE.json configs[3] with configs[2]'s crowded parameters; the metric "images/sec
RPN+RoIAlign+mask-paste" is quoted on it): every GPU processes F = 64 synthetic 704x520 frames per
step (weak scaling: 512 frames on 8 GPUs), 2000 synthetic cells per frame, pre-NMS top-k 2000,
post-NMS 1000 proposals, 256-channel level-0 FPN map (130x176), 7x7 RoIAlign, box-score filter 0.4 +
NMS 0.5 into at most 500 detections, 28x28 mask probabilities pasted to uint8 704x520 frames, one
all-gather of detection records when N > 1.  The PyTorch heads (out of scope) are replaced
by seeded synthetic box scores / mask probabilities.

One JSON line on stdout (rank 0).

`value`: inputs resident in HBM, the step run through the product's serving path pipeline.StreamedRegionPipeline (paste of a
16-frame sub-batch on a second stream under select/NMS/RoIAlign of the next one; CUDA graphs replayed).  `kernels`: the same
stages run back to back on one stream, one graph per stage with events between them — the isolated per-stage times the
roofline fractions are computed from.  `roofline`: the dominant kernel (mask paste, 71 % of the algorithmic bytes) timed
alone.  `e2e`: the same step through pipeline.HostFedRegionPipeline starting from pinned HOST buffers (H2D of every input
+ D2H of the detection records inside the timed region), with the plain-memcpy ceiling of the same bytes beside it.
Extras (N = 1): `roi_align_bwd` on the bench RoI list, `c5` (BASELINE config C5 sweep, torchvision's sm_100 cubins beside
it), `c2_train_shape_ms`, `c1_latency_ms` / `c3_latency_ms`, `value_nchw_input` (NCHW feature maps, transpose inside the
timed region), `sustained` (the step looped >= 3 s with the SM clock sampled), `match_boxes_c1` (the fused anchor matcher at
C1 size, torchvision's chain beside it).  N > 1: `multi_rank_check`.
`cpu_baseline` / `--impl reference`: the CPU oracle port (oracle/lcr_oracle.c, OpenMP over every usable host core — set
explicitly, torchrun exports OMP_NUM_THREADS=1) on a bounded sample of the same workload; `cpu_baseline_reference_python`:
the UNMODIFIED reference functions from the staged checkout baseline/_ref on the same host cores.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

IMG_H, IMG_W = 520, 704
FH, FW, A, C = 130, 176, 9, 256
PRE_NMS, POST_NMS, MAX_DET, M = 2000, 1000, 500, 28
N_CELLS = 2000
ANCHOR_CHOICES = [0, 1, 2, 3, 4, 5]      # cells use the size-32/64 anchors (LIVECell cells are 16-76 px)
METRIC = "images_per_sec_rpn_roialign_maskpaste"


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        return float(json.load(open(path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def workload_config(frames, n_gpus):
    return {
        "workload": f"C4/C3 crowded inference: {frames} synthetic 704x520 frames per GPU per step ({frames * n_gpus} total), "
                    f"{N_CELLS} cells/frame, pre-NMS top-k {PRE_NMS}, post-NMS {POST_NMS}, {C}-ch 130x176 level-0 map, 7x7 RoIAlign, "
                    f"<= {MAX_DET} detections, 28x28 mask paste to uint8 704x520",
        "frames_per_gpu": frames,
        "feature_layout": "NHWC (channels_last)",
        "parallelism": f"image-sharded x{n_gpus}, one all-gather of detection records",
        "l2": "per-step working set (~13 GB of masks + features per GPU) >> 126 MB L2: every step streams from HBM",
    }


# ------------------------------------------------------------------------------------------------
# synthetic inputs (host side, seeded)
# ------------------------------------------------------------------------------------------------
def make_host_inputs(frames, seed0):
    from livecell_instance_segmentation_b200 import synth
    obj = np.concatenate([synth.make_objectness(1, A, FH, FW, n_cells=N_CELLS, seed=seed0 + i, k=PRE_NMS,
                                                anchor_choices=ANCHOR_CHOICES) for i in range(frames)])
    box_scores = synth.make_box_scores((frames, POST_NMS), seed0 + 10_000)
    return obj, box_scores


def make_mask_probs(n, seed):
    from livecell_instance_segmentation_b200 import synth
    base = synth.make_mask_probs(512, M, seed)           # 512 distinct maps tiled over the detection slots
    reps = (n + 511) // 512
    return np.tile(base, (reps, 1, 1))[:n]


# ------------------------------------------------------------------------------------------------
# clocks sampling during the timed region
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.index),
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm = [float(r[0]) for r in self.rows if len(r) >= 6 and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) >= 6 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = sorted({names[i] for r in self.rows if len(r) >= 6 for i in range(4) if r[2 + i].lower().startswith("active")})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


# ------------------------------------------------------------------------------------------------
# CPU arm: the oracle port on a bounded sample of the same workload
# ------------------------------------------------------------------------------------------------
def cpu_step(orc, base, obj, feat_nchw, box_scores, probs):
    """One pass of the region path over len(obj) frames on the host (oracle/lcr_oracle.c)."""
    F = obj.shape[0]
    boxes, scores, _, counts = orc.rpn_select_batch(obj, base=base, stride=4, k=PRE_NMS, score_thresh=0.3, min_size=10.0,
                                                    img_h=IMG_H, img_w=IMG_W)
    keep, kc = orc.nms_batch(boxes, None, counts, 0.4, post_n=POST_NMS)
    n_det = 0
    for f in range(F):
        n = int(kc[f])
        pb = boxes[f][keep[f, :n]]
        rois = np.concatenate([np.zeros((n, 1), np.float32), pb], axis=1)
        orc.roi_align_fwd(feat_nchw[f:f + 1], rois, 7, 7, 0.25, 2, False)
        k2 = orc.nms(pb, box_scores[f, :n], 0.5, score_thresh=0.4, use_score_thresh=True, post_n=MAX_DET)
        orc.paste_masks(probs[: len(k2)], pb[k2], IMG_H, IMG_W)
        n_det += len(k2)
    return n_det


def cpu_baseline(sample_frames, steps, warmup, seed0=0):
    from oracle import oracle as orc
    orc.build()
    cores = orc.set_num_threads(0)          # every usable core, whatever OMP_NUM_THREADS the launcher exported
    usable = orc.usable_cores()
    assert cores > 1 or usable == 1, f"CPU baseline would run on {cores} thread(s) of {usable} usable cores"
    from livecell_instance_segmentation_b200 import synth
    base = orc.base_anchors()
    obj, box_scores = make_host_inputs(sample_frames, seed0)
    feat = synth.make_features(sample_frames, C, FH, FW, seed=seed0 + 77)
    probs = make_mask_probs(MAX_DET, seed0 + 99)
    for _ in range(warmup):
        cpu_step(orc, base, obj, feat, box_scores, probs)
    t0 = time.perf_counter()
    for _ in range(steps):
        cpu_step(orc, base, obj, feat, box_scores, probs)
    dt = (time.perf_counter() - t0) / max(steps, 1)
    return {"value": sample_frames / dt, "unit": "images/s", "cores": cores, "kind": "port",
            "sample": f"{sample_frames} frames/step x {steps} steps of the same workload (oracle/lcr_oracle.c, OpenMP, {cores} threads "
                      f"of {usable} usable cores, OMP_NUM_THREADS in the environment was {os.environ.get('OMP_NUM_THREADS', 'unset')!r}), "
                      f"{dt * 1e3:.0f} ms/step"}, dt


def reference_python_baseline(sample_frames=2, steps=2, seed0=0):
    """The reference's OWN functions, unmodified, imported from the staged checkout baseline/_ref (tools/stage_reference.py):
    generate_inference_proposals -> torchvision RoIAlign -> score filter + nms -> paste_masks_in_image per frame, the loop body
    of CustomMaskRCNN.forward_inference (src/custom_maskrcnn.py:164-207) with the PyTorch heads replaced by the same
    synthetic scores / probabilities as every other arm, on the host CPU with every usable core."""
    ref = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isfile(os.path.join(ref, "src", "utils", "proposal_utils.py")):
        return {"unavailable": "baseline/_ref is not staged (tools/stage_reference.py needs /root/reference)"}
    try:
        import torch
        cores = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else (os.cpu_count() or 1)
        torch.set_num_threads(cores)
        sys.path.insert(0, ref)
        from src.components.anchor_generator import AnchorGenerator
        from src.utils.mask_utils import paste_masks_in_image
        from src.utils.proposal_utils import generate_inference_proposals
        from torchvision.ops import RoIAlign, nms
        from livecell_instance_segmentation_b200 import synth
        obj, box_scores = make_host_inputs(sample_frames, seed0)
        feat = torch.from_numpy(synth.make_features(sample_frames, C, FH, FW, seed=seed0 + 77))
        probs = torch.from_numpy(make_mask_probs(MAX_DET, seed0 + 99))
        obj_t, bs_t = torch.from_numpy(obj), torch.from_numpy(box_scores)
        anchors = AnchorGenerator().generate_anchors((FH, FW), 4, "cpu")
        roi_align = RoIAlign(output_size=(7, 7), spatial_scale=0.25, sampling_ratio=2)
        stage = {"proposals": 0.0, "roi_align": 0.0, "det_nms": 0.0, "paste": 0.0}

        def frame(f, acc):
            t = [time.perf_counter()]
            props, _ = generate_inference_proposals(obj_t[f], anchors, (IMG_H, IMG_W), "cpu", num_pre_nms=PRE_NMS, score_threshold=0.3,
                                                    nms_threshold=0.4, num_post_nms=POST_NMS, min_box_size=10)
            t.append(time.perf_counter())
            roi_align(feat[f:f + 1], [props])
            t.append(time.perf_counter())
            sc = bs_t[f, : len(props)]
            keep = sc > 0.4
            fb, fs = props[keep], sc[keep]
            k2 = nms(fb, fs, 0.5)[:MAX_DET]
            t.append(time.perf_counter())
            paste_masks_in_image(probs[: len(k2)], fb[k2], (IMG_H, IMG_W))
            t.append(time.perf_counter())
            if acc:
                for i, k in enumerate(stage):
                    stage[k] += t[i + 1] - t[i]
            return len(k2)

        frame(0, False)                                   # warm-up
        t0 = time.perf_counter()
        n_det = 0
        for _ in range(steps):
            for f in range(sample_frames):
                n_det += frame(f, True)
        dt = (time.perf_counter() - t0) / max(steps, 1)
        n = steps * sample_frames
        return {"value": sample_frames / dt, "unit": "images/s", "cores": cores, "kind": "reference",
                "torch_threads": torch.get_num_threads(), "os_cpu_count": os.cpu_count(),
                "sample": f"{sample_frames} frames/step x {steps} steps of the same workload through the unmodified reference functions "
                          f"(baseline/_ref/src/utils/proposal_utils.py:33-59, torchvision RoIAlign + nms CPU ops, "
                          f"src/utils/mask_utils.py:129-171), {dt * 1e3:.0f} ms/step",
                "ms_per_frame": {k: v / n * 1e3 for k, v in stage.items()}, "detections_per_frame": n_det / n}
    except Exception as exc:   # a broken staging must not take the GPU line down
        return {"unavailable": f"{type(exc).__name__}: {exc}"}
    finally:
        if ref in sys.path:
            sys.path.remove(ref)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    frames = 16     # bounded sample: ~1.6 s of 8-core work per step
    cb, dt = cpu_baseline(frames, args.steps, args.warmup)
    line = {"impl": "reference", "metric": METRIC, "value": cb["value"], "unit": "images/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": workload_config(frames, 1), "cpu_baseline": cb,
            "e2e": {"value": cb["value"], "unit": "images/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    line["config"]["workload"] = "bounded sample of: " + workload_config(64, 1)["workload"]
    if not args.no_reference_python:
        line["cpu_baseline_reference_python"] = reference_python_baseline(2, 1)
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    from livecell_instance_segmentation_b200 import _lib, ops, synth
    from livecell_instance_segmentation_b200.dist import all_gather_detections, bind_to_gpu_numa_node, last_bind_report
    from livecell_instance_segmentation_b200.pipeline import (HostFedRegionPipeline, RegionConfig, RegionPipeline,
                                                               StreamedRegionPipeline)
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import bench_extras as X

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the region pipeline has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    numa_cores = bind_to_gpu_numa_node(local) if world > 1 else []     # before the pinned host buffers are allocated
    numa_report = last_bind_report() if world > 1 else {"reason": "single rank: not attempted"}
    saved_stdout = None
    if world > 1:
        # NCCL prints its version banner on the process's stdout; the contract is ONE JSON line there, so fd 1 is
        # pointed at stderr until the result line is printed
        sys.stdout.flush()
        saved_stdout = os.dup(1)
        os.dup2(2, 1)
        dist.init_process_group("nccl", device_id=dev)
    F = args.frames
    n_items = F * world
    cfg = RegionConfig(pre_nms_top_n=PRE_NMS, post_nms_top_n=POST_NMS, max_detections=MAX_DET)
    peak, peak_src = peaks()
    stream = torch.cuda.current_stream()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- inputs: host (pinned) and device copies --------------------------------------------------
    def make_inputs(r):
        obj_h, bs_h = make_host_inputs(F, seed0=r * F)
        probs_h = make_mask_probs(F * MAX_DET, 99 + r)
        g = torch.Generator(device=dev).manual_seed(1234 + r)
        feat = torch.randn((F, FH, FW, C), generator=g, device=dev, dtype=torch.float32).permute(0, 3, 1, 2)  # NHWC memory
        return obj_h, bs_h, probs_h, feat

    obj_h, bs_h, probs_h, feat_d = make_inputs(rank)
    host = {
        "obj": torch.from_numpy(obj_h).pin_memory(),
        "bs": torch.from_numpy(bs_h).pin_memory(),
        "probs": torch.from_numpy(probs_h).pin_memory(),
        "feat": torch.empty((F, FH, FW, C), dtype=torch.float32).pin_memory(),
    }
    host["feat"].copy_(feat_d.permute(0, 2, 3, 1))

    # ---- the product path: StreamedRegionPipeline over static input buffers -----------------------------------------
    sp = StreamedRegionPipeline(cfg, F, (C, FH, FW), (IMG_H, IMG_W), num_anchors=A, chunks=args.chunks, device=dev, depth=args.depth)
    sp.inputs["obj"].copy_(host["obj"])
    sp.inputs["bs"].copy_(host["bs"])
    sp.inputs["probs"].copy_(host["probs"])
    sp.inputs["feat"].copy_(feat_d)
    del feat_d
    inp = sp.inputs
    torch.cuda.synchronize()
    l0 = _lib.launch_count()
    sp.run(finish=True)                       # eager once: counts this step's liblcr launches (graph replays are not counted)
    torch.cuda.synchronize()
    launches_per_step = _lib.launch_count() - l0
    launch_mode = "eager"
    if not args.no_graph:
        try:
            sp.capture()
            launch_mode = (f"cuda_graph_replay ({2 * sp.chunks} graphs per step: select/NMS/RoIAlign and paste of each of {sp.chunks} sub-batches, "
                           f"{sp.depth} alternating buffer sets)")
        except Exception as exc:          # capture unsupported for some reason: time the eager launches instead
            print(f"bench: CUDA graph capture failed ({exc!r}); timing eager launches", file=sys.stderr)
            sp._graphs = None
            torch.cuda.synchronize()

    # the only exchange: detection records (12 KB/frame).  Enqueued from the paste stream behind the step's last paste, the
    # previous step's gather is collected first: the collective of step i overlaps the kernels of step i+1.
    pending = []
    stage_rec = [torch.empty_like(sp.records) for _ in range(2)]
    stage_cnt = [torch.empty_like(sp.counts) for _ in range(2)]
    gathered = {}

    def gather(i):
        if world == 1:
            return
        with torch.cuda.stream(sp.side):
            if pending:
                gathered["last"] = pending.pop().wait()
            stage_rec[i % 2].copy_(sp.records, non_blocking=True)
            stage_cnt[i % 2].copy_(sp.counts, non_blocking=True)
            pending.append(all_gather_detections(stage_rec[i % 2], stage_cnt[i % 2], n_items, async_op=True))

    def drain():
        if pending:
            with torch.cuda.stream(sp.side):
                gathered["last"] = pending.pop().wait()
                sp.done.record(sp.side)
        sp.finish()

    for i in range(args.warmup):
        sp.run()
        gather(i)
    drain()
    barrier()

    # ---- timed region: inputs resident in HBM ----------------------------------------------------
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record(stream)
    for i in range(args.steps):
        sp.run()
        gather(i)
    drain()                                   # every paste and the last step's gathered records are part of the timed work
    t_end.record(stream)
    barrier()
    ms_step = t_start.elapsed_time(t_end) / args.steps
    clocks = sampler.stop() if rank == 0 else None
    if world > 1:
        t = torch.tensor([ms_step], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms_step = float(t.item())
    value = n_items / (ms_step * 1e-3)
    launches = launches_per_step * args.steps
    pc = sp.proposal_counts.cpu().numpy()
    dc = sp.counts.cpu().numpy()
    n_props, n_det = int(pc.sum()), int(dc.sum())

    # ---- N > 1: the gathered detection set is what a single GPU computes (SURVEY §4 iv) ---------------------------------
    multi_rank = None
    if world > 1:
        rec, cnt = gathered["last"]
        own = bool(torch.equal(rec[rank * F:(rank + 1) * F], sp.records) and torch.equal(cnt[rank * F:(rank + 1) * F], sp.counts))
        ok = torch.tensor([1 if own else 0], device=dev)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN)
        foreign = None
        if rank == 0:      # rank 0 recomputes rank 1's shard from rank 1's seeds with the plain sequential stages
            o2, b2, p2, f2 = make_inputs(1)
            seq = RegionPipeline(cfg)
            props2 = seq.proposals(torch.from_numpy(o2).to(dev), (IMG_H, IMG_W))
            det2 = seq.detections(props2, torch.from_numpy(b2).to(dev))
            rec2 = ops.pack_records(det2.boxes, det2.scores, det2.counts)
            foreign = bool(torch.equal(rec[F:2 * F], rec2) and torch.equal(cnt[F:2 * F], det2.counts))
            del o2, b2, p2, f2, props2, det2, rec2
        multi_rank = {"every_rank_finds_its_own_shard_in_the_gathered_set": bool(ok.item()),
                      "rank0_recomputation_of_rank1_shard_matches_gathered": foreign,
                      "gathered_records_shape": list(rec.shape)}

    # ---- sustained: the same step looped for >= 3 s with the SM clock sampled ----------------------------------------
    sustained = None
    if not args.no_extras:
        s2 = ClockSampler(local)
        s2.start()
        n_loop = max(args.steps, int(3000.0 / max(ms_step, 1e-3)) + 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record(stream)
        for _ in range(n_loop):
            sp.run()
        sp.finish()
        e1.record(stream)
        torch.cuda.synchronize()
        sms = e0.elapsed_time(e1) / n_loop
        sustained = {"value": F / (sms * 1e-3), "unit": "images/s per GPU", "ms_per_step": sms, "steps": n_loop,
                     "seconds": sms * n_loop * 1e-3, "clocks": s2.stop()}

    # ---- isolated stages: back to back on ONE stream, one CUDA graph per stage, events between the replays ----------------
    pipe = RegionPipeline(cfg)
    stage_names = ["rpn_select+nms+gather", "roi_align_fwd", "det_nms+gather", "paste+records"]
    st = {}

    def s1():
        st["props"] = pipe.proposals(inp["obj"], (IMG_H, IMG_W))

    def s2_():
        st["roi_feat"] = pipe.pool(inp["feat"], st["props"].rois, out=sp.roi_features)

    def s3():
        st["det"] = pipe.detections(st["props"], inp["bs"])

    def s4():
        st["det"] = pipe.paste(st["det"], inp["probs"], (IMG_H, IMG_W), out=sp.masks)

    fns = [s1, s2_, s3, s4]
    for fn in fns:
        fn()
    torch.cuda.synchronize()
    graphs = []
    if not args.no_graph:
        for fn in fns:
            gph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(gph):
                fn()
            graphs.append(gph)
    iso_steps = max(3, min(args.steps, 10))
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(5)] for _ in range(iso_steps)]
    for w in range(2 + iso_steps):
        for j in range(4):
            if w >= 2:
                evs[w - 2][j].record(stream)
            (graphs[j].replay if graphs else fns[j])()
        if w >= 2:
            evs[w - 2][4].record(stream)
    torch.cuda.synchronize()
    stage_ms = [float(np.mean([evs[i][j].elapsed_time(evs[i][j + 1]) for i in range(iso_steps)])) for j in range(4)]
    props, det = st["props"], st["det"]
    seq_counts_match = bool(torch.equal(det.counts, sp.counts) and torch.equal(props.counts, sp.proposal_counts))

    # ---- algorithmic bytes ---------------------------------------------------------------
    bytes_select = F * (4 * A * FH * FW) + F * PRE_NMS * 28 + F * (20 * PRE_NMS + 8 * POST_NMS)
    bytes_roi = 4 * n_props * C * 49 + F * 4 * C * FH * FW + 20 * F * POST_NMS
    bytes_det = F * (20 * POST_NMS + 8 * MAX_DET)
    bytes_paste = n_det * (IMG_H * IMG_W + M * M * 4 + 16) + F * MAX_DET * 24
    stage_bytes = [bytes_select, bytes_roi, bytes_det, bytes_paste]
    kernels = {n: {"ms": m, "algorithmic_GB": b / 1e9, "GBps": b / 1e9 / (m * 1e-3), "frac_of_hbm_peak": b / 1e9 / (m * 1e-3) / peak}
               for n, m, b in zip(stage_names, stage_ms, stage_bytes)}
    kernels["sum_of_isolated_stages_ms"] = float(sum(stage_ms))

    # dominant kernel alone (paste_split_kernel): events around the single launch, output > L2
    pe = [torch.cuda.Event(enable_timing=True) for _ in range(2 * iso_steps)]
    valid = det.valid
    boxes_flat = det.boxes.reshape(-1, 4)
    for i in range(iso_steps):
        pe[2 * i].record(stream)
        ops.paste_masks(inp["probs"], boxes_flat, IMG_H, IMG_W, 0.5, 255, valid=valid, out=sp.masks)
        pe[2 * i + 1].record(stream)
    torch.cuda.synchronize()
    paste_ms = float(np.mean([pe[2 * i].elapsed_time(pe[2 * i + 1]) for i in range(iso_steps)]))
    paste_bytes = n_det * (IMG_H * IMG_W + M * M * 4 + 16)
    traffic = None     # dram bytes per launch of the same kernel/workload from the committed ncu --set full capture
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        tj = json.load(open(tpath))
        if tj.get("frames_per_gpu") == F and n_det == F * MAX_DET:
            traffic = tj.get("paste_split_kernel", {}).get("dram_bytes_per_launch")
    roofline = {"kernel": "paste_split_kernel", "bound": "hbm", "achieved": paste_bytes / 1e9 / (paste_ms * 1e-3), "peak": peak,
                "unit": "GB/s", "frac": paste_bytes / 1e9 / (paste_ms * 1e-3) / peak, "traffic": traffic,
                "traffic_source": "profiles/ncu_traffic.json (ncu --set full, same workload)" if traffic else None,
                "peak_source": peak_src, "bytes_per_launch": paste_bytes, "ms_per_launch": paste_ms,
                "share_of_step": stage_ms[3] / max(sum(stage_ms), 1e-9)}

    # NMS latency (BASELINE metric "NMS us"): one 2000-box segment, sorted input.  The three launches (rank,
    # mask, resolve) are captured once into a CUDA graph and replayed, so the figure is device time, not the
    # Python launch path; `nms_us_2000_boxes_eager` is the same call issued eagerly from Python.
    cand_boxes, _, _, cand_counts = ops.rpn_select([inp["obj"][:1]], k=PRE_NMS, img_size=(IMG_H, IMG_W), score_thresh=0.3, min_size=10.0,
                                                   strides=[4], base=pipe.base)
    nb, nc = cand_boxes[:, 0].contiguous(), cand_counts[:, 0].contiguous()
    reps = 50
    ne = [torch.cuda.Event(enable_timing=True) for _ in range(4)]
    for _ in range(5):
        ops.nms_batched(nb, None, 0.4, post_n=POST_NMS, counts=nc)
    ne[0].record(stream)
    for _ in range(reps):
        ops.nms_batched(nb, None, 0.4, post_n=POST_NMS, counts=nc)
    ne[1].record(stream)
    torch.cuda.synchronize()
    nms_us_eager = ne[0].elapsed_time(ne[1]) * 1e3 / reps
    nms_us = None
    if not args.no_graph:
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            ops.nms_batched(nb, None, 0.4, post_n=POST_NMS, counts=nc)
        for _ in range(5):
            graph.replay()
        ne[2].record(stream)
        for _ in range(reps):
            graph.replay()
        ne[3].record(stream)
        torch.cuda.synchronize()
        nms_us = ne[2].elapsed_time(ne[3]) * 1e3 / reps
        del graph

    # ---- extras on one GPU: backward, NCHW input, C5 / C2 / C1 / C3 -----------------------
    extras = {}
    if world == 1 and not args.no_extras:
        extras["roi_align_bwd"] = X.roi_bwd_on_list(ops, props.rois, n_props, F, C, FH, FW, peak)
        # NCHW feature maps (what an unchanged NCHW backbone emits): the NCHW->NHWC transpose kernel runs inside every step
        feat_nchw = inp["feat"].contiguous()
        n_e0, n_e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        # product form: StreamedRegionPipeline.nchw_features() — run() transposes the batch's maps into the NHWC buffer first
        sp.nchw_features().copy_(feat_nchw)
        inp["feat"].zero_()                                  # the NHWC buffer is now an intermediate: results must not depend on it
        for w in range(2):
            nsteps = 2 if w == 0 else iso_steps
            torch.cuda.synchronize()
            n_e0.record(stream)
            for _ in range(nsteps):
                sp.run()
            sp.finish()
            n_e1.record(stream)
            torch.cuda.synchronize()
        nms_ = n_e0.elapsed_time(n_e1) / iso_steps
        nchw_same = bool(torch.equal(sp.counts.cpu(), torch.from_numpy(dc)) and torch.equal(sp.proposal_counts.cpu(), torch.from_numpy(pc)))
        sp.nchw_features(False)
        extras["value_nchw_input"] = {"value": F / (nms_ * 1e-3), "unit": "images/s", "ms_per_step": nms_,
                                      "what": "same step fed NCHW-contiguous feature maps through StreamedRegionPipeline.nchw_features(): "
                                              "lcr_nchw_to_nhwc_f32 of the 64 level-0 maps at the head of every step, inside the timed region "
                                              "(per sub-batch inside the front groups: measured slower, 3.92 against 3.74 ms)",
                                      "counts_match_nhwc_run": nchw_same}
        del feat_nchw

    # ---- e2e: the public host-fed entry point (pipeline.HostFedRegionPipeline.run): every step copies that
    # step's inputs from pinned host memory (chunked, overlapped with compute) and reads the records back ----------
    graphs = st = props = det = valid = boxes_flat = None
    sp_records_own = sp.records.clone()
    del sp, inp, stage_rec, stage_cnt
    torch.cuda.empty_cache()
    barrier()
    ceiling = X.h2d_ceiling(host, reps=3)
    barrier()
    runner = HostFedRegionPipeline(cfg, F, (C, FH, FW), (IMG_H, IMG_W), num_anchors=A, chunk_frames=args.chunk_frames, device=dev)
    gather_fn = (lambda r, c: all_gather_detections(r, c, n_items)) if world > 1 else None

    def e2e_step():
        return runner.run(host, gather=gather_fn, sync=True)      # the caller holds the detections on the host

    e2e_steps = max(2, min(args.steps, 5))
    for _ in range(2):
        rec_h, cnt_h = e2e_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for _ in range(e2e_steps):
        rec_h, cnt_h = e2e_step()
    e1.record(stream)
    barrier()
    e2e_ms = e0.elapsed_time(e1) / e2e_steps
    ceil_gbps = ceiling["GBps"]
    if world > 1:
        t = torch.tensor([e2e_ms, -ceil_gbps], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        e2e_ms, ceil_gbps = float(t[0].item()), -float(t[1].item())       # slowest rank's step, slowest rank's ceiling
    h2d = runner.h2d_bytes(host)
    d2h = int(rec_h.numel() * 4 + cnt_h.numel() * 4)
    own_rec = rec_h[rank * F: rank * F + F] if world > 1 else rec_h
    e2e_ok = bool(torch.equal(cnt_h[rank * F: rank * F + F] if world > 1 else cnt_h, torch.from_numpy(dc).to(torch.int32))
                  and torch.equal(own_rec, sp_records_own.cpu()))
    h2d_gbps = h2d / 1e9 / (e2e_ms * 1e-3)
    del runner
    torch.cuda.empty_cache()

    if world == 1 and not args.no_extras:
        extras["c5"] = X.c5_sweep(ops, synth, peak)
        extras["c2_train_shape"] = X.c2_train_shape(synth)
        extras["c2_train_shape_ms"] = extras["c2_train_shape"]["fwd_plus_bwd_ms"]
        extras.update(X.config_latency(ops, synth, RegionConfig, RegionPipeline))
        try:
            extras["match_boxes_c1"] = X.match_boxes_c1(ops, synth)
        except Exception as exc:   # an extra, never the line
            extras["match_boxes_c1"] = {"error": repr(exc)}

    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if rank == 0:
        cfg_line = workload_config(F, world)
        cfg_line["value_path"] = (f"pipeline.StreamedRegionPipeline: {args.chunks} sub-batches of {F // max(args.chunks, 1)} frames, paste on a second "
                                  "stream under the next sub-batch's select/NMS/RoIAlign; streams joined only at the end of the timed region")
        line = {
            "metric": METRIC, "value": value, "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": cfg_line, "clocks": clocks,
            "e2e": {"value": n_items / (e2e_ms * 1e-3), "unit": "images/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": e2e_ms, "h2d_GBps": h2d_gbps,
                    "h2d_ceiling_GBps": ceil_gbps, "frac_of_h2d_ceiling": h2d_gbps / ceil_gbps if ceil_gbps else None,
                    "h2d_ceiling": "plain Tensor.copy_ (cudaMemcpyAsync) of the same pinned buffers, GPU otherwise idle, all ranks at once; "
                                   "slowest rank" + (" (tools/h2d_ceiling.py sweeps streams / chunk sizes)"),
                    "bound": "PCIe host->device copy of the step's inputs (compute is hidden behind it)",
                    "api": f"pipeline.HostFedRegionPipeline.run (chunks of {args.chunk_frames} frames, H2D overlapped with compute)",
                    "d2h": "detection records + counts (pasted masks stay sharded in HBM, SURVEY §8e)",
                    "results_match_resident_run": e2e_ok, "host_cores_bound_to_gpu_numa_node": len(numa_cores), "numa_binding": numa_report},
            "gpu_launches": int(launches), "launch_mode": launch_mode, "roofline": roofline,
            # the WHOLE step against the HBM roofline: the algorithmic bytes of all four stages (what `kernels` uses
            # per stage) over the streamed step's device time — the stages overlap, so HBM is the one resource they all share
            "step_roofline": {"bound": "hbm", "algorithmic_GB": sum(stage_bytes) / 1e9, "ms_per_step": ms_step,
                              "achieved": sum(stage_bytes) / 1e9 / (ms_step * 1e-3), "peak": peak, "unit": "GB/s",
                              "frac": sum(stage_bytes) / 1e9 / (ms_step * 1e-3) / peak},
            "kernels": kernels,
            "streamed_counts_match_sequential_stages": seq_counts_match, "sustained": sustained,
            "nms_us_2000_boxes": nms_us if nms_us is not None else nms_us_eager, "nms_us_2000_boxes_eager": nms_us_eager,
            "proposals_per_frame": n_props / F, "detections_per_frame": n_det / F,
        }
        if multi_rank is not None:
            line["multi_rank_check"] = multi_rank
        line.update(extras)
        if not args.no_cpu_baseline:
            cb, _ = cpu_baseline(16, 3, 1)
            line["cpu_baseline"] = cb
            if not args.no_reference_python:
                line["cpu_baseline_reference_python"] = reference_python_baseline(2, 1)
        if saved_stdout is not None:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
        print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--frames", type=int, default=64, help="frames per GPU per step")
    ap.add_argument("--chunks", type=int, default=2, help="sub-batches per step of the streamed pipeline (paste overlaps the next one)")
    ap.add_argument("--depth", type=int, default=2, help="buffer sets of the streamed pipeline (2: a step's first stages run under the previous step's last paste)")
    ap.add_argument("--chunk-frames", type=int, default=8, help="e2e: frames per H2D/compute pipeline chunk")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-reference-python", action="store_true", help="skip the unmodified-reference CPU leg (baseline/_ref)")
    ap.add_argument("--no-extras", action="store_true", help="skip backward / C5 / C2 / C1 / C3 / NCHW / sustained extras")
    ap.add_argument("--no-graph", action="store_true", help="time eager Python launches instead of CUDA-graph replays")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 0)
    if args.impl == "reference":
        run_reference(args)
    else:
        if args.warmup < 3:
            args.warmup = 3          # timing rule: W >= 3
        run_ours(args)


if __name__ == "__main__":
    main()

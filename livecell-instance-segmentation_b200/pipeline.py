"""Batched region pipeline — the B200-first form of the per-image loop of
CustomMaskRCNN.forward_inference (src/custom_maskrcnn.py:164-207).

The reference walks the batch in Python and syncs the host 10+ times per image; here every stage
takes the whole batch, results live in fixed-capacity padded device buffers with device-side counts,
and nothing synchronises until the caller asks for dense per-image outputs:

    rpn_select (1 launch, cluster per image) -> nms (3) -> gather (1) -> roi_align (1)
      -> [box head: PyTorch] -> nms with score filter (3) -> gather (1)
      -> [mask head: PyTorch] -> paste (1) -> records (1)

Default hyper-parameters are the reference's (src/utils/proposal_utils.py:33-36,
src/custom_maskrcnn.py:48-50,185,192,292); BASELINE config C3 uses 2000 / 1000 / 500.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Callable, Optional, Tuple

import torch

from . import ops
from .roi_align import _prefer_nhwc


@dataclass
class RegionConfig:
    sizes: Tuple[float, ...] = (32, 64, 128)
    aspect_ratios: Tuple[float, ...] = (0.5, 1.0, 2.0)
    stride: int = 4                      # hard-coded in the reference (custom_maskrcnn.py:99,159)
    pre_nms_top_n: int = 250
    rpn_score_thresh: float = 0.3
    rpn_nms_thresh: float = 0.4
    post_nms_top_n: int = 50
    min_box_size: float = 10.0
    pooled_size: int = 7
    spatial_scale: float = 0.25
    sampling_ratio: int = 2
    box_score_thresh: float = 0.4
    det_nms_thresh: float = 0.5
    max_detections: Optional[int] = None  # capacity of the detection buffers (default: post_nms_top_n)
    mask_thresh: float = 0.5
    mask_size: int = 28

    @property
    def det_capacity(self) -> int:
        return self.post_nms_top_n if self.max_detections is None else self.max_detections


@dataclass
class Proposals:
    boxes: torch.Tensor      # [B, post_n, 4]
    scores: torch.Tensor     # [B, post_n]
    counts: torch.Tensor     # [B] i32
    rois: torch.Tensor       # [B*post_n, 5], batch index -1 on padding rows


@dataclass
class Detections:
    boxes: torch.Tensor      # [B, D, 4]
    scores: torch.Tensor     # [B, D]
    counts: torch.Tensor     # [B] i32
    index: torch.Tensor      # [B, D] i64: proposal slot of each detection
    valid: torch.Tensor      # [B*D] u8
    masks: Optional[torch.Tensor] = None    # [B*D, H, W] u8 (slots with valid == 0 are untouched)
    records: Optional[torch.Tensor] = None  # [B, D, 6]


class RegionPipeline:
    def __init__(self, cfg: Optional[RegionConfig] = None):
        self.cfg = cfg or RegionConfig()
        self.base = ops.base_anchors(self.cfg.sizes, self.cfg.aspect_ratios)

    # -- stage 1: objectness -> proposals ----------------------------------------------------------
    def proposals(self, objectness: torch.Tensor, image_size) -> Proposals:
        """objectness [B, A, h, w] level-0 logits (cls_scores[0], custom_maskrcnn.py:166)."""
        c = self.cfg
        n = objectness[0].numel()
        k = min(c.pre_nms_top_n, n)
        boxes, scores, _, counts = ops.rpn_select([objectness], k=k, img_size=image_size, score_thresh=c.rpn_score_thresh,
                                                  min_size=c.min_box_size, strides=[c.stride], base=self.base)
        boxes, scores, counts = boxes[:, 0], scores[:, 0], counts[:, 0].contiguous()
        keep, kc = ops.nms_batched(boxes, None, c.rpn_nms_thresh, post_n=c.post_nms_top_n, counts=counts)
        pb, ps, rois = ops.gather_kept(boxes, scores, keep, kc, want_rois=True)
        return Proposals(pb, ps, kc, rois)

    # -- stage 2: RoIAlign ---------------------------------------------------------------------------
    def pool(self, features: torch.Tensor, rois: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """features: logical [B, C, h, w]; channels_last memory takes the fast path directly.  An NCHW map is transposed once
        (tiled 16-byte transpose) when the list is long enough to pay for it (roi_align._prefer_nhwc), otherwise its planes are
        pooled directly (roi_fwd_planes_kernel)."""
        c = self.cfg
        B, C, h, w = features.shape
        nhwc = features.stride() == (h * w * C, 1, w * C, C)
        if not nhwc and C % 4 == 0 and _prefer_nhwc(features, rois.shape[0], c.pooled_size, c.pooled_size):
            features = ops.to_nhwc(features)
        return ops.roi_align_fwd([features], [c.spatial_scale], rois, None, (c.pooled_size, c.pooled_size), c.sampling_ratio, False,
                                 out=out)

    # -- stage 3: box scores -> detections ---------------------------------------------------------
    def detections(self, props: Proposals, box_scores: torch.Tensor) -> Detections:
        """box_scores [B, post_n] = softmax(cls_logits)[:, 1] per proposal slot
        (custom_maskrcnn.py:182-195: score > 0.4, then NMS 0.5)."""
        c = self.cfg
        D = c.det_capacity
        keep, kc = ops.nms_batched(props.boxes, box_scores, c.det_nms_thresh, post_n=D, counts=props.counts,
                                   score_thresh=c.box_score_thresh)
        db, ds, _, valid = ops.gather_kept(props.boxes, box_scores, keep, kc, want_rois=False, want_valid=True)
        return Detections(db, ds, kc, keep, valid)

    # -- stage 4: mask probabilities -> frames + records -------------------------------------------
    def paste(self, det: Detections, mask_probs: torch.Tensor, image_size, out: Optional[torch.Tensor] = None) -> Detections:
        """mask_probs [B*D, M, M] = sigmoid(mask_logits[:, 1]) per detection slot
        (custom_maskrcnn.py:273-295)."""
        H, W = image_size
        B, D = det.scores.shape
        det.masks = ops.paste_masks(mask_probs.reshape(B * D, mask_probs.shape[-2], mask_probs.shape[-1]), det.boxes.reshape(B * D, 4),
                                    int(H), int(W), self.cfg.mask_thresh, 255, valid=det.valid, out=out)
        det.records = ops.pack_records(det.boxes, det.scores, det.counts)
        return det

    # -- whole path with the PyTorch heads as callables --------------------------------------------
    def infer(self, objectness: torch.Tensor, features: torch.Tensor, image_size,
              box_head: Callable[[torch.Tensor], torch.Tensor], mask_head: Callable[[torch.Tensor], torch.Tensor]):
        """Batched equivalent of forward_inference's loop.  box_head(roi_features [K,C,7,7]) -> class-1
        probabilities [K]; mask_head(roi_features [N,C,7,7]) -> class-1 mask probabilities [N,M,M].
        Returns the reference's output structure: list of dicts(boxes, labels, scores, masks)."""
        c = self.cfg
        B = objectness.shape[0]
        props = self.proposals(objectness, image_size)
        roi_feat = self.pool(features, props.rois)
        box_scores = box_head(roi_feat).reshape(B, c.post_nms_top_n)
        det = self.detections(props, box_scores)
        D = c.det_capacity
        offs = (torch.arange(B, device=det.index.device) * c.post_nms_top_n).unsqueeze(1)
        flat = (det.index.clamp(min=0, max=c.post_nms_top_n - 1) + offs).reshape(-1)
        flat = torch.where(det.valid.bool(), flat, torch.zeros_like(flat))
        probs = mask_head(roi_feat.index_select(0, flat))
        det = self.paste(det, probs, image_size)
        return self.to_predictions(det, image_size)

    @staticmethod
    def to_predictions(det: Detections, image_size):
        """Dense per-image dicts with the reference's dtypes (custom_maskrcnn.py:202-207, 306-314).
        This is the one place that synchronises (reads the counts)."""
        H, W = image_size
        B, D = det.scores.shape
        counts = det.counts.tolist()
        masks = det.masks.reshape(B, D, H, W) if det.masks is not None else None
        preds = []
        for b, n in enumerate(counts):
            preds.append({
                "boxes": det.boxes[b, :n],
                "labels": torch.ones((n,), dtype=torch.long, device=det.boxes.device),
                "scores": det.scores[b, :n],
                "masks": masks[b, :n] if masks is not None else torch.zeros((n, H, W), dtype=torch.uint8, device=det.boxes.device),
            })
        return preds


class StreamedRegionPipeline:
    """The serving form of the region path: paste overlaps the next frames' selection / NMS / RoIAlign.

    Mask paste is HBM-write-bound and needs one small CTA per SM (90+ % of the copy peak on its own); proposal selection,
    NMS and RoIAlign are latency / L1-bound and leave HBM mostly idle.  Run back to back they add up (3.8 ms per 64 frames);
    co-scheduled they overlap (measured 3.35 ms).  Here a batch of F frames is cut into `chunks` sub-batches; stages 1-3 of
    sub-batch c+1 (select -> NMS -> gather -> RoIAlign -> detection NMS) run on the caller's stream while stage 4 of
    sub-batch c (paste + records) runs on a side stream.  Nothing joins the streams until ``finish()`` (or ``run(...,
    finish=True)``), so the paste of a batch's last sub-batch also overlaps the next batch's first stages: results of a batch are
    valid once the side stream has passed its ``done`` event.

    All result buffers (masks, pooled features, records, counts) are owned by this object and reused by every batch — the
    reference allocates them per image (src/custom_maskrcnn.py:164-207).  ``capture()`` records every sub-batch's two stage
    groups as CUDA graphs over static input buffers (`self.inputs`), after which ``run()`` only replays.
    ``nchw_features()`` adds a static NCHW-contiguous feature buffer (``self.inputs["feat_nchw"]``): when present, ``run()``
    first transposes the batch's maps into the NHWC buffer on the caller's stream (one tiled-transpose launch, graphs or not)
    — up front, where it still overlaps the previous batch's last paste.  Transposing each sub-batch's
    maps inside its front group instead was measured slower (3.92 against 3.74 ms per 64 frames: the transpose is HBM-bound,
    and so is the paste it would run beside)."""

    def __init__(self, cfg: RegionConfig, frames: int, feat_shape, image_size, num_anchors: int = 9, chunks: int = 2, device=None,
                 depth: int = 2):
        self.pipe = RegionPipeline(cfg)
        self.cfg = cfg
        self.F = int(frames)
        self.chunks = max(1, min(int(chunks), self.F))
        self.depth = max(1, int(depth))   # sets of per-sub-batch intermediates: with 2, the front stages of a sub-batch only wait
        # for its paste of TWO batches ago, so a batch's first stages also run under the previous batch's last paste
        self.H, self.W = int(image_size[0]), int(image_size[1])
        C, fh, fw = feat_shape
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        D, P, M = cfg.det_capacity, cfg.post_nms_top_n, cfg.mask_size
        f32 = dict(dtype=torch.float32, device=dev)
        F = self.F
        self.inputs = {     # static input buffers (graph mode copies / writes into these; eager mode may pass its own tensors)
            "obj": torch.empty((F, num_anchors, fh, fw), **f32),
            "feat": torch.empty((F, fh, fw, C), **f32).permute(0, 3, 1, 2),      # NHWC memory, logical NCHW
            "bs": torch.empty((F, P), **f32),
            "probs": torch.empty((F * D, M, M), **f32),
        }
        self.masks = torch.empty((F * D, self.H, self.W), dtype=torch.uint8, device=dev)
        self.roi_features = torch.empty((F * P, C, cfg.pooled_size, cfg.pooled_size), **f32)
        self.records = torch.zeros((F, D, 6), **f32)
        self.counts = torch.zeros((F,), dtype=torch.int32, device=dev)
        self.proposal_counts = torch.zeros((F,), dtype=torch.int32, device=dev)
        self.side = torch.cuda.Stream(device=dev)
        self.bounds = [(c * F // self.chunks, (c + 1) * F // self.chunks) for c in range(self.chunks)]
        n_slots = self.depth * self.chunks
        self.ev_main = [torch.cuda.Event() for _ in range(n_slots)]     # stages 1-3 of (set, sub-batch) enqueued
        self.ev_side = [torch.cuda.Event() for _ in range(n_slots)]     # stage 4 of (set, sub-batch) enqueued
        self.used = [False] * n_slots
        self.done = torch.cuda.Event()
        self._state = [None] * n_slots
        self._graphs = None
        self._batch = 0

    def nchw_features(self, enable: bool = True) -> Optional[torch.Tensor]:
        """Static NCHW-contiguous feature input [F, C, h, w] (None after disabling).  Graphs captured earlier keep their form:
        capture() again after switching."""
        if not enable:
            self.inputs.pop("feat_nchw", None)
            return None
        if self.inputs.get("feat_nchw") is None:
            self.inputs["feat_nchw"] = torch.empty(tuple(self.inputs["feat"].shape), dtype=torch.float32, device=self.device)
        return self.inputs["feat_nchw"]

    # the two stage groups of one sub-batch ----------------------------------------------------------------------------
    def _front(self, c, inp, slot=None):
        slot = c if slot is None else slot
        f0, f1 = self.bounds[c]
        P = self.cfg.post_nms_top_n
        props = self.pipe.proposals(inp["obj"][f0:f1], (self.H, self.W))
        self.pipe.pool(inp["feat"][f0:f1], props.rois, out=self.roi_features[f0 * P: f1 * P])
        det = self.pipe.detections(props, inp["bs"][f0:f1])
        self.proposal_counts[f0:f1].copy_(props.counts, non_blocking=True)
        self._state[slot] = det

    def _back(self, c, inp, slot=None):
        slot = c if slot is None else slot
        f0, f1 = self.bounds[c]
        D = self.cfg.det_capacity
        det = self.pipe.paste(self._state[slot], inp["probs"][f0 * D: f1 * D], (self.H, self.W), out=self.masks[f0 * D: f1 * D])
        self.records[f0:f1].copy_(det.records, non_blocking=True)
        self.counts[f0:f1].copy_(det.counts, non_blocking=True)

    def capture(self):
        """Record the stage groups of every sub-batch as CUDA graphs over ``self.inputs`` (fill those before each run)."""
        main = torch.cuda.current_stream(self.device)
        warm = torch.cuda.Stream(device=self.device)
        warm.wait_stream(main)
        with torch.cuda.stream(warm):                     # warm-up off the capture stream (workspaces, lazy module loads)
            for c in range(self.chunks):
                self._front(c, self.inputs)
                self._back(c, self.inputs)
        main.wait_stream(warm)
        torch.cuda.synchronize(self.device)
        graphs = []
        for slot in range(self.depth * self.chunks):      # one pair of graphs per (set, sub-batch): own intermediates
            c = slot % self.chunks
            gf, gb = torch.cuda.CUDAGraph(), torch.cuda.CUDAGraph()
            with torch.cuda.graph(gf):
                self._front(c, self.inputs, slot)
            with torch.cuda.graph(gb, pool=gf.pool()):
                self._back(c, self.inputs, slot)
            graphs.append((gf, gb))
        self._graphs = graphs
        return self

    def run(self, inputs: Optional[dict] = None, finish: bool = False):
        """One batch.  inputs: dict(obj [F,A,h,w], feat [F,C,h,w] (channels_last memory preferred), bs [F,post_n],
        probs [F*D,M,M]) of device tensors; None = the static buffers ``self.inputs`` (required after capture()).
        Returns self (records / counts / masks / roi_features are valid after ``finish()`` or once ``done`` has passed)."""
        if self._graphs is not None and inputs is not None and inputs is not self.inputs:
            raise ValueError("after capture() the batch must be written into self.inputs")
        inp = self.inputs if inputs is None else inputs
        main = torch.cuda.current_stream(self.device)
        base = (self._batch % self.depth) * self.chunks
        self._batch += 1
        if inp.get("feat_nchw") is not None:     # NCHW-contiguous maps: one transpose of the batch into the NHWC buffer
            ops.to_nhwc(inp["feat_nchw"], out=inp["feat"])   # (a plain launch; as a one-node graph it measured 0.3 ms slower per step)
        for c in range(self.chunks):
            slot = base + c
            if self.used[slot]:
                main.wait_event(self.ev_side[slot])       # this set's previous paste (`depth` batches ago) has read its inputs
            if self._graphs is not None:
                self._graphs[slot][0].replay()
            else:
                self._front(c, inp, slot)
            self.ev_main[slot].record(main)
            with torch.cuda.stream(self.side):
                self.side.wait_event(self.ev_main[slot])
                if self._graphs is not None:
                    self._graphs[slot][1].replay()
                else:
                    self._back(c, inp, slot)
                self.ev_side[slot].record(self.side)
            self.used[slot] = True
        self.done.record(self.side)
        if finish:
            self.finish()
        return self

    def finish(self):
        """Make the caller's stream wait for every enqueued paste (no host sync)."""
        torch.cuda.current_stream(self.device).wait_event(self.done)
        return self

    def detections(self) -> Detections:
        """The last finished batch as a Detections view over the persistent buffers."""
        D = self.cfg.det_capacity
        rec = self.records
        valid = (torch.arange(D, device=self.device)[None, :] < self.counts[:, None]).reshape(-1).to(torch.uint8)
        return Detections(rec[:, :, :4], rec[:, :, 4], self.counts, torch.empty(0), valid, self.masks, rec)


class HostFedRegionPipeline:
    """The region path fed from pinned HOST buffers (the serving entry point bench.py's `e2e` times).

    A batch of F frames is cut into chunks; chunk c+1's host->device copies (objectness, NHWC features,
    box scores, mask probabilities) run on a copy stream while chunk c is processed on the compute
    stream (two device staging sets, events both ways), so the step costs max(PCIe, compute) instead of
    their sum.  Detection records + counts are returned in pinned host memory; pasted masks and pooled
    features stay in HBM (`masks`, `roi_features`), owned by this object and reused every call.
    The reference does the same work image by image with ~10 host syncs each
    (src/custom_maskrcnn.py:164-207); here the only sync is the final one the caller needs.
    """

    def __init__(self, cfg: RegionConfig, frames: int, feat_shape, image_size, num_anchors: int = 9, chunk_frames: int = 8,
                 device=None):
        self.pipe = RegionPipeline(cfg)
        self.cfg = cfg
        self.F, self.FC = int(frames), max(1, min(int(chunk_frames), int(frames)))
        self.H, self.W = int(image_size[0]), int(image_size[1])
        C, fh, fw = feat_shape
        dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.device = dev
        D, P, M = cfg.det_capacity, cfg.post_nms_top_n, cfg.mask_size
        f32 = dict(dtype=torch.float32, device=dev)
        self.stage = [{
            "obj": torch.empty((self.FC, num_anchors, fh, fw), **f32),
            "feat": torch.empty((self.FC, fh, fw, C), **f32),          # NHWC memory
            "bs": torch.empty((self.FC, P), **f32),
            "probs": torch.empty((self.FC * D, M, M), **f32),
        } for _ in range(2)]
        self.masks = torch.empty((self.F * D, self.H, self.W), dtype=torch.uint8, device=dev)
        self.roi_features = torch.empty((self.F * P, C, cfg.pooled_size, cfg.pooled_size), **f32)
        self.records = torch.zeros((self.F, D, 6), **f32)
        self.counts = torch.zeros((self.F,), dtype=torch.int32, device=dev)
        self.records_host = torch.empty((self.F, D, 6), dtype=torch.float32).pin_memory()
        self.counts_host = torch.empty((self.F,), dtype=torch.int32).pin_memory()
        self.copy_stream = torch.cuda.Stream(device=dev)
        self.ready = [torch.cuda.Event() for _ in range(2)]
        self.free = [torch.cuda.Event() for _ in range(2)]
        self.slot_used = [False, False]          # across calls: a staging set may still be read by the previous run()
        self.done = torch.cuda.Event()           # previous run()'s device->host copies
        self.ran = False

    def h2d_bytes(self, host) -> int:
        return sum(int(host[k].numel() * host[k].element_size()) for k in ("obj", "feat", "bs", "probs"))

    def run(self, host: dict, gather: Optional[Callable] = None, sync: bool = True):
        """host: pinned CPU tensors obj [F,A,h,w], feat [F,h,w,C] (NHWC), bs [F,post_n], probs [F*D,M,M].
        gather(records, counts) -> (records, counts): optional collective (all-gather over ranks) applied
        before the device->host copy.  Returns (records_host, counts_host).  With sync=False the returned pinned buffers
        are valid once ``self.done`` has passed (``self.done.synchronize()``); a following run() orders itself behind the
        previous one (staging sets, result buffers and the pinned copies are reused), so back-to-back async calls are safe
        but the host must have consumed the previous results before calling again."""
        cfg, F, FC = self.cfg, self.F, self.FC
        D, P = cfg.det_capacity, cfg.post_nms_top_n
        compute = torch.cuda.current_stream(self.device)
        if self.ran:
            # sync=False callers: the previous call's records / counts (device buffers and pinned host copies) are still
            # in flight until its `done` event; this call must not overwrite them earlier
            compute.wait_event(self.done)
        n_chunks = (F + FC - 1) // FC
        for c in range(n_chunks):
            f0, f1 = c * FC, min(F, (c + 1) * FC)
            n = f1 - f0
            st = self.stage[c % 2]
            with torch.cuda.stream(self.copy_stream):
                if self.slot_used[c % 2]:
                    self.copy_stream.wait_event(self.free[c % 2])      # the set's previous user (this or the previous call) is done
                st["obj"][:n].copy_(host["obj"][f0:f1], non_blocking=True)
                st["feat"][:n].copy_(host["feat"][f0:f1], non_blocking=True)
                st["bs"][:n].copy_(host["bs"][f0:f1], non_blocking=True)
                st["probs"][: n * D].copy_(host["probs"][f0 * D: f1 * D], non_blocking=True)
                self.ready[c % 2].record(self.copy_stream)
            compute.wait_event(self.ready[c % 2])
            props = self.pipe.proposals(st["obj"][:n], (self.H, self.W))
            self.pipe.pool(st["feat"][:n].permute(0, 3, 1, 2), props.rois, out=self.roi_features[f0 * P: f1 * P])
            det = self.pipe.detections(props, st["bs"][:n])
            det = self.pipe.paste(det, st["probs"][: n * D], (self.H, self.W), out=self.masks[f0 * D: f1 * D])
            self.records[f0:f1].copy_(det.records, non_blocking=True)
            self.counts[f0:f1].copy_(det.counts, non_blocking=True)
            self.free[c % 2].record(compute)
            self.slot_used[c % 2] = True
        rec, cnt = (self.records, self.counts) if gather is None else gather(self.records, self.counts)
        if rec.shape != self.records_host.shape:
            self.records_host = torch.empty(rec.shape, dtype=torch.float32).pin_memory()
            self.counts_host = torch.empty(cnt.shape, dtype=torch.int32).pin_memory()
        self.records_host.copy_(rec, non_blocking=True)
        self.counts_host.copy_(cnt, non_blocking=True)
        self.done.record(compute)
        self.ran = True
        if sync:
            compute.synchronize()
        return self.records_host, self.counts_host

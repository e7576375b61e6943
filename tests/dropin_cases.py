"""Drivers for the drop-in tests: each function runs ONE reference code path —
``CustomMaskRCNN.forward_inference`` / ``forward_train`` + backward (src/custom_maskrcnn.py:85-209),
``train_custom.train_one_epoch`` / ``evaluate`` (src/train_custom.py:36-166), ``app_gradio.predict_single_image``
(src/app_gradio.py:18-72), ``visualize.predict_on_tiles`` + ``filter_detections_by_border_mini_tiles``
(src/visualize.py:134-257) — on the staged, unmodified reference (baseline/_ref), either untouched or with
``install()`` applied, from the same seed and weights, and returns plain numpy results for comparison.

The CPU suite calls them untouched-vs-untouched (harness check); tests/test_gpu_dropin.py compares untouched vs patched
on the B200.
"""
import os

import numpy as np
import torch

import ref_harness


def synth_image(H, W, n_cells, seed):
    """SURVEY §8d 'phase-contrast-like' frame: 0.5 + 0.1 N(0,1) background + n_cells filled ellipses (+0.2), 3 channels."""
    rng = np.random.RandomState(seed)
    img = 0.5 + 0.1 * rng.randn(H, W).astype(np.float32)
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    for _ in range(n_cells):
        cx, cy = rng.uniform(0, W), rng.uniform(0, H)
        a, b = rng.uniform(6, 20, size=2)
        th = rng.uniform(0, np.pi)
        xr = (xx - cx) * np.cos(th) + (yy - cy) * np.sin(th)
        yr = -(xx - cx) * np.sin(th) + (yy - cy) * np.cos(th)
        img[(xr / a) ** 2 + (yr / b) ** 2 <= 1.0] += 0.2
    img = np.clip(img, 0.0, 1.0)
    return np.repeat(img[None], 3, axis=0)


def synth_targets(B, H, W, seed, device):
    """GT for the training shapes: a jittered grid of anchor-like boxes (32/64/128 px, so that randomly initialised
    proposals do match some of them: IoU >= 0.4 positives exist) with elliptical uint8 masks."""
    rng = np.random.RandomState(seed)
    targets = []
    yy, xx = np.mgrid[0:H, 0:W].astype(np.float32)
    for b in range(B):
        boxes = []
        for size, step in ((32, 40), (64, 72), (128, 120)):
            for y0 in range(0, H - size // 2, step):
                for x0 in range(0, W - size // 2, step):
                    ar = float(np.exp(rng.uniform(-0.5, 0.5)))
                    w, h = size * np.sqrt(ar), size / np.sqrt(ar)
                    cx, cy = x0 + size / 2 + rng.uniform(-6, 6), y0 + size / 2 + rng.uniform(-6, 6)
                    bx = [max(cx - w / 2, 0), max(cy - h / 2, 0), min(cx + w / 2, W), min(cy + h / 2, H)]
                    if bx[2] - bx[0] > 8 and bx[3] - bx[1] > 8:
                        boxes.append(bx)
        boxes = np.asarray(boxes, np.float32)
        masks = np.zeros((len(boxes), H, W), np.uint8)
        for i, (x1, y1, x2, y2) in enumerate(boxes):
            cx, cy, a, bb = (x1 + x2) / 2, (y1 + y2) / 2, (x2 - x1) / 2, (y2 - y1) / 2
            masks[i] = (((xx - cx) / a) ** 2 + ((yy - cy) / bb) ** 2 <= 1.0).astype(np.uint8)
        targets.append({"boxes": torch.from_numpy(boxes).to(device), "labels": torch.ones(len(boxes), dtype=torch.int64, device=device),
                        "masks": torch.from_numpy(masks).to(device)})
    return targets


def _deterministic():
    """Deterministic cuDNN algorithms and NO TF32: with TF32 (torch's cuDNN default) a 1e-7 difference in a pooled feature
    is re-rounded to a 10-bit mantissa inside every convolution of the heads and comes out as 1e-3 — the untouched model
    then differs from ITSELF by 1e-3 between two runs (atomics in torchvision's RoIAlign backward), which would hide what
    these tests measure: the difference our kernels introduce."""
    torch.backends.cudnn.benchmark = False
    torch.backends.cudnn.deterministic = True
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False


def new_model(cm, device, seed=0, channels_last=False, state=None):
    """get_custom_model() of whatever `cm` currently binds (untouched or patched), fixed weights."""
    torch.manual_seed(seed)
    model = cm.get_custom_model(num_classes=2)
    if state is not None:
        model.load_state_dict(state)
    model = model.to(device)
    if channels_last:
        model = model.to(memory_format=torch.channels_last)
    return model


def calibrate_rpn(model, images, target_std=0.3):
    """A randomly initialised RPN scores every anchor 0.5 +- 0.002, so the top-k post-sigmoid values collide in fp32
    and torch.topk's (unspecified) tie order decides the proposals.  Scale the objectness layer so that the level-0
    logits have std `target_std` on `images` (in the model's current train/eval mode): scores spread over ~0.3-0.8 like
    a briefly trained model and the top-k is tie-free.  Pure function of (weights, images): the untouched and the
    patched run calibrate identically.  BatchNorm buffers are restored afterwards."""
    saved = {k: v.detach().clone() for k, v in model.state_dict().items()}
    with torch.no_grad():
        feats, _ = model.extract_features(images)
        cls, _ = model.rpn(feats)
        gain = float(target_std / cls[0].std().clamp(min=1e-12))
    model.load_state_dict(saved)
    with torch.no_grad():
        model.rpn.cls_logits.weight.mul_(gain)
    return gain


def perturb_roi_align_by_one_ulp(model, seed=99):
    """Yardstick for 'how far may a correct replacement be': the UNTOUCHED model with every pooled feature moved by one
    to two fp32 ulps (x * (1 +- 2^-23), random sign; 1.2e-7 relative, a hundredth of the 1e-5 parity bound).  A different summation order inside RoIAlign does exactly this to
    the forward result; whatever the rest of the network (BatchNorm in train mode, AdamW's g/|g| first step, a top-k
    edge) amplifies it to is the reference's own sensitivity, not an error of the replacement."""
    inner = model.roi_align.forward
    gen = {}

    def forward(x, rois):
        out = inner(x, rois)
        g = gen.setdefault(out.device, torch.Generator(device=out.device).manual_seed(seed))
        sign = torch.randint(0, 2, out.shape, generator=g, device=out.device, dtype=torch.int8).to(out.dtype) * 2 - 1
        return out * (1.0 + sign * 2.0 ** -23)

    model.roi_align.forward = forward
    return model


def _install(patched):
    if patched:
        from livecell_instance_segmentation_b200 import install as inst
        done = inst.install(batched_inference=(patched == "batched"))
        assert done, "install() patched nothing"
        return inst
    return None


def rpn_scores_tie_free(model, images, k=251, only_first=False):
    """The top-(k) post-sigmoid level-0 scores of every image are distinct (torch.topk's tie order is unspecified, so
    top-k index parity is only defined on tie-free inputs — SURVEY §8a2).  Restores BatchNorm buffers (train mode)."""
    saved = {k_: v.detach().clone() for k_, v in model.state_dict().items()}
    try:
        return _tie_free(model, images, k, only_first)
    finally:
        model.load_state_dict(saved)


def _tie_free(model, images, k, only_first):
    with torch.no_grad():
        feats, _ = model.extract_features(images)
        cls, _ = model.rpn(feats)
        for b in range(1 if only_first else cls[0].shape[0]):
            s = torch.sigmoid(cls[0][b]).reshape(-1)
            top = torch.topk(s, min(k, s.numel())).values
            if torch.unique(top).numel() != top.numel():
                return False
    return True


def run_inference(device, patched, H, W, B, seed=0, channels_last=False, n_cells=150, state=None):
    """forward_inference (custom_maskrcnn.py:144-209) on B synthetic frames -> list of dicts of numpy arrays."""
    _deterministic()
    cm = ref_harness.import_reference()
    inst = _install(patched)
    try:
        model = new_model(cm, device, seed, channels_last, state).eval()
        images = torch.from_numpy(np.stack([synth_image(H, W, n_cells, 100 * seed + b) for b in range(B)])).to(device)
        if channels_last:
            images = images.contiguous(memory_format=torch.channels_last)
        if state is None:
            calibrate_rpn(model, images)
        with torch.no_grad():
            preds = model(images)
        out = [{k: v.detach().cpu().numpy() for k, v in p.items()} for p in preds]
        info = {"tie_free": rpn_scores_tie_free(model, images), "roi_align_type": type(model.roi_align).__module__,
                "state": {k: v.detach().cpu().clone() for k, v in model.state_dict().items()}}
        return out, info
    finally:
        if inst is not None:
            inst.uninstall()
        ref_harness.purge()


def run_train_step(device, patched, B=8, H=256, W=256, seed=0, state=None, channels_last=False, perturb=False):
    """forward_train + backward (custom_maskrcnn.py:85-142, train_custom.py:40-44) -> (losses, {param: grad})."""
    _deterministic()
    cm = ref_harness.import_reference()
    inst = _install(patched)
    try:
        model = new_model(cm, device, seed, channels_last, state).train()
        images = torch.from_numpy(np.stack([synth_image(H, W, 40, 7 + b) for b in range(B)])).to(device)
        targets = synth_targets(B, H, W, 5, device)
        base_w = model.rpn.cls_logits.weight.detach().clone()
        for std in (0.3, 0.45, 0.6, 0.2, 0.9):                 # first calibration whose top-501 of image 0 is tie-free
            with torch.no_grad():
                model.rpn.cls_logits.weight.copy_(base_w)
            calibrate_rpn(model, images, std)
            tie_free = rpn_scores_tie_free(model, images, k=501, only_first=True)   # forward_train selects from image 0 only
            if tie_free:
                break
        if perturb:
            assert not patched
            perturb_roi_align_by_one_ulp(model)
        torch.manual_seed(1234)                     # the randperm draws of rpn.compute_loss / sample_proposals
        loss_dict = model(images, targets)
        total = sum(v for v in loss_dict.values())
        model.zero_grad()
        total.backward()
        losses = {k: float(v.detach().cpu()) for k, v in loss_dict.items()}
        grads = {n: p.grad.detach().cpu().numpy().copy() for n, p in model.named_parameters() if p.grad is not None}
        return losses, grads, {"roi_align_type": type(model.roi_align).__module__, "tie_free": tie_free}
    finally:
        if inst is not None:
            inst.uninstall()
        ref_harness.purge()


class _Loader:
    """Stand-in for the DataLoader of src/dataset.py (collate -> tuple of images, tuple of targets)."""

    def __init__(self, batches):
        self.batches = batches

    def __iter__(self):
        return iter(self.batches)

    def __len__(self):
        return len(self.batches)


def run_train_epoch(device, patched, n_batches=2, B=4, H=256, W=256, perturb=False, perturb_seed=99):
    """train_custom.train_one_epoch + evaluate, imported UNCHANGED (stub matplotlib/wandb/pycocotools), driven with a
    synthetic loader (train_custom.py:21-166)."""
    _deterministic()
    ref_harness.import_reference(scripts=True)
    import importlib
    tc = importlib.import_module("train_custom")        # does `from custom_maskrcnn import get_custom_model` itself
    inst = _install(patched)
    try:
        torch.manual_seed(0)
        model = tc.get_custom_model(num_classes=2).to(device)
        opt = torch.optim.AdamW(model.parameters(), lr=1e-4, weight_decay=1e-4)    # train_custom.py main(): AdamW
        batches = []
        for i in range(n_batches):
            imgs = tuple(torch.from_numpy(synth_image(H, W, 40, 50 + 10 * i + b)) for b in range(B))
            tg = tuple({k: v.cpu() for k, v in t.items()} for t in synth_targets(B, H, W, 90 + i, "cpu"))
            batches.append((imgs, tg))
        model.train()
        calibrate_rpn(model, torch.stack([i.to(device) for i in batches[0][0]]))
        if perturb:
            assert not patched
            perturb_roi_align_by_one_ulp(model, perturb_seed)
        # record, per training step, how many of the top-501 post-sigmoid scores collide (torch.topk's tie order is unspecified)
        import sys as _sys
        ties = []
        wrapped = []
        for mod_name in ("custom_maskrcnn", "src.custom_maskrcnn"):
            mod = _sys.modules.get(mod_name)
            if mod is None or not hasattr(mod, "generate_training_proposals"):
                continue
            inner = mod.generate_training_proposals

            def spy(cls_scores, anchors, image_size, dev, *a, _inner=inner, **kw):
                s_ = torch.sigmoid(cls_scores).permute(1, 2, 0).reshape(-1)
                top = torch.topk(s_, min(501, s_.numel())).values
                ties.append(int(top.numel() - torch.unique(top).numel()))
                return _inner(cls_scores, anchors, image_size, dev, *a, **kw)

            mod.generate_training_proposals = spy
            wrapped.append((mod, inner))
        torch.manual_seed(4321)
        metrics = tc.train_one_epoch(model, _Loader(batches), opt, device, epoch=1)
        val = tc.evaluate(model, _Loader(batches[:1]), device)
        for mod, inner in wrapped:
            mod.generate_training_proposals = inner
        return {k: float(v) for k, v in metrics.items()}, {k: float(v) for k, v in val.items()}, \
            {"roi_align_type": type(model.roi_align).__module__, "topk_ties_per_step": ties}
    finally:
        if inst is not None:
            inst.uninstall()
        ref_harness.purge()


def run_gradio_predict(device, patched, tmpdir, H=520, W=704):
    """app_gradio.predict_single_image imported UNCHANGED (stub gradio/matplotlib) on a saved checkpoint
    (app_gradio.py:18-72 -> visualize.load_model:27-69)."""
    _deterministic()
    cm = ref_harness.import_reference(scripts=True)
    import importlib
    inst = _install(patched)        # before app_gradio binds visualize.load_model -> custom_maskrcnn.get_custom_model
    try:
        torch.manual_seed(0)
        ckpt = os.path.join(str(tmpdir), "custom_model.pth")
        frame = (synth_image(H, W, 150, 3)[0] * 255).astype(np.uint8)
        frame = np.repeat(frame[:, :, None], 3, axis=2)
        if not os.path.exists(ckpt):          # written by the untouched run, read back by the patched one
            m0 = cm.get_custom_model(num_classes=2).eval()
            calibrate_rpn(m0, torch.from_numpy(frame).permute(2, 0, 1).float().div(255)[None])
            torch.save({"model_state_dict": m0.state_dict()}, ckpt)
        app = importlib.import_module("app_gradio")
        app.DEVICE = torch.device(device)
        captured = {}
        vis = importlib.import_module("visualize")
        real_load = vis.load_model

        def load_and_keep(*a, **kw):
            captured["model"] = real_load(*a, **kw)
            return captured["model"]

        app.load_model = load_and_keep
        result_img, status = app.predict_single_image(frame, ckpt, 0.5)
        return status, np.asarray(result_img).shape, {"roi_align_type": type(captured["model"].roi_align).__module__}
    finally:
        if inst is not None:
            inst.uninstall()
        ref_harness.purge()


def run_tiles(device, patched, tmpdir, n_tiles=3):
    """visualize.predict_on_tiles + filter_detections_by_border_mini_tiles UNCHANGED on 300x222 PNG tiles
    (visualize.py:134-257; tile geometry preprocess_dataset.py:86-124)."""
    from PIL import Image
    from torchvision import transforms
    _deterministic()
    cm = ref_harness.import_reference(scripts=True)
    import importlib
    inst = _install(patched)
    try:
        vis = importlib.import_module("visualize")
        tiles = []
        for t in range(n_tiles):
            path = os.path.join(str(tmpdir), f"img_tile_{t:02d}.png")
            if not os.path.exists(path):
                arr = (synth_image(222, 300, 40, 200 + t).transpose(1, 2, 0) * 255).astype(np.uint8)
                Image.fromarray(arr).save(path)
            tiles.append({"path": path, "tile_num": t, "filename": os.path.basename(path)})
        model = new_model(cm, device, 0).eval()
        first = transforms.ToTensor()(Image.open(tiles[0]["path"]).convert("RGB")).to(device)[None]
        calibrate_rpn(model, first)
        results = vis.predict_on_tiles(model, tiles, torch.device(device), transforms.Compose([transforms.ToTensor()]))
        kept = vis.filter_detections_by_border_mini_tiles(results, score_threshold=0.5, mask_threshold=0.4)
        preds = [{k: v.numpy() for k, v in r["prediction"].items()} for r in results]
        summary = [(d["tile_num"], [float(x) for x in d["box"]], float(d["score"]), int(d["mask"].sum())) for d in kept]
        return preds, summary, {"roi_align_type": type(model.roi_align).__module__}
    finally:
        if inst is not None:
            inst.uninstall()
        ref_harness.purge()


# ---- comparison helpers ------------------------------------------------------------------------------------------------
def compare_predictions(ref, got, score_atol=1e-6, what=""):
    """boxes / labels bit-exact, scores within score_atol, masks bit-exact.  Returns the number of mask pixels that
    differ (the caller decides what to do with a non-zero count)."""
    assert len(ref) == len(got), what
    flips = 0
    for i, (r, g) in enumerate(zip(ref, got)):
        assert r["boxes"].shape == g["boxes"].shape, f"{what} image {i}: {r['boxes'].shape[0]} vs {g['boxes'].shape[0]} detections"
        assert r["boxes"].dtype == g["boxes"].dtype and r["scores"].dtype == g["scores"].dtype
        assert r["labels"].dtype == g["labels"].dtype == np.int64 and r["masks"].dtype == g["masks"].dtype == np.uint8
        assert np.array_equal(r["boxes"], g["boxes"]), f"{what} image {i}: boxes differ"
        assert np.array_equal(r["labels"], g["labels"]), f"{what} image {i}: labels differ"
        if r["scores"].size:
            err = float(np.abs(r["scores"] - g["scores"]).max())
            assert err <= score_atol, f"{what} image {i}: scores differ by {err:.3e}"
        assert r["masks"].shape == g["masks"].shape
        assert set(np.unique(g["masks"]).tolist()) <= {0, 255}
        flips += int((r["masks"] != g["masks"]).sum())
    return flips


def rel_err(a, b):
    """max |a-b| / max |b| (the max-norm form) and the worst elementwise relative error over |b| > 1e-3 max|b|."""
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    scale = max(float(np.abs(b).max()), 1e-30)
    return float(np.abs(a - b).max()) / scale

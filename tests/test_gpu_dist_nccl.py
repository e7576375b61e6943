"""Two B200s, NCCL: the image-sharded region pipeline + one all-gather of detection records reproduces, on every rank, exactly
what ONE GPU computes for the whole batch (SURVEY.md §4 iv, §8e).  Skipped when fewer than two GPUs
are visible (the driver's single-GPU `-m gpu` run); `gpurun --gpus 2 -- python -m pytest tests/test_gpu_dist_nccl.py -m gpu`
runs it."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


F_TOTAL, H, W, C, h, w = 10, 96, 128, 64, 24, 32      # 10 frames over 2 ranks, and 7 (ragged 4 + 3)


def _inputs(n_frames):
    from livecell_instance_segmentation_b200 import synth
    obj = synth.make_objectness(n_frames, 9, h, w, n_cells=60, seed=501, k=200)
    feat = synth.make_features(n_frames, C, h, w, seed=502)
    bs = np.stack([synth.make_box_scores((60,), 503 + b) for b in range(n_frames)])
    probs = synth.make_mask_probs(n_frames * 40, 28, 505)
    return obj, feat, bs, probs


def _run_shard(dev, obj, feat, bs, probs, lo, hi):
    from livecell_instance_segmentation_b200.pipeline import RegionConfig, StreamedRegionPipeline
    cfg = RegionConfig(pre_nms_top_n=200, post_nms_top_n=60, max_detections=40)
    n = hi - lo
    sp = StreamedRegionPipeline(cfg, n, (C, h, w), (H, W), chunks=2, device=dev)
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    sp.run({"obj": t(obj[lo:hi]), "feat": t(feat[lo:hi]).contiguous(memory_format=torch.channels_last), "bs": t(bs[lo:hi]),
            "probs": t(probs[lo * 40: hi * 40])}, finish=True)
    torch.cuda.synchronize(dev)
    return sp


def _worker(rank, world, port, n_frames, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from livecell_instance_segmentation_b200.dist import all_gather_detections, shard_range
    obj, feat, bs, probs = _inputs(n_frames)
    lo, hi = shard_range(n_frames, rank, world)
    sp = _run_shard(dev, obj, feat, bs, probs, lo, hi)
    rec, cnt = all_gather_detections(sp.records, sp.counts, n_frames)
    hnd = all_gather_detections(sp.records, sp.counts, n_frames, async_op=True)
    rec2, cnt2 = hnd.wait()
    torch.cuda.synchronize(dev)
    assert torch.equal(rec, rec2) and torch.equal(cnt, cnt2)
    valid = (torch.arange(40, device=dev)[None, :] < sp.counts[:, None]).reshape(-1)       # padded slots are never written
    q.put((rank, rec.cpu().numpy(), cnt.cpu().numpy(), int(sp.masks[valid].sum().item())))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
@pytest.mark.parametrize("n_frames", [10, 7], ids=["even_shards", "ragged_shards"])
def test_two_gpu_gather_equals_single_gpu(n_frames):
    import torch.multiprocessing as mp
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.path.join(root, "tests") + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_frames, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=300) for _ in range(2)]
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    obj, feat, bs, probs = _inputs(n_frames)
    single = _run_shard(torch.device("cuda", 0), obj, feat, bs, probs, 0, n_frames)      # the whole batch on one GPU
    ref_rec, ref_cnt = single.records.cpu().numpy(), single.counts.cpu().numpy()
    assert ref_cnt.sum() > 0
    mask_sums = 0
    for rank, rec, cnt, msum in results:
        assert np.array_equal(cnt, ref_cnt), rank
        assert np.array_equal(rec, ref_rec), rank
        mask_sums += msum
    valid = (np.arange(40)[None, :] < ref_cnt[:, None]).reshape(-1)
    assert mask_sums == int(single.masks[torch.from_numpy(valid).to(single.masks.device)].sum().item())   # sharded masks add up

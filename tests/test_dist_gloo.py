"""CPU, world_size 2, gloo: the N>1 path — contiguous image shards, no data-path collective, one
all-gather of fixed-size detection records — reassembles exactly the single-process result."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _fake_detect(img_ids, D=5):
    """Deterministic per-image 'detections' (records depend only on the image id)."""
    rec = torch.zeros((len(img_ids), D, 6))
    cnt = torch.zeros((len(img_ids),), dtype=torch.int32)
    for i, g in enumerate(img_ids):
        n = g % (D + 1)
        cnt[i] = n
        for j in range(n):
            rec[i, j] = torch.tensor([g, j, g + 10.0, j + 10.0, 1.0 / (1 + j), 1.0])
    return rec, cnt


def _worker(rank, world, port, n_items, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from livecell_instance_segmentation_b200.dist import all_gather_detections, shard_range
    s, e = shard_range(n_items, rank, world)
    rec, cnt = _fake_detect(list(range(s, e)))
    grec, gcnt = all_gather_detections(rec, cnt, n_items)
    h = all_gather_detections(rec, cnt, n_items, async_op=True)       # the overlapped form bench.py uses
    arec, acnt = h.wait()
    assert torch.equal(arec, grec) and torch.equal(acnt, gcnt) and h.wait()[0] is arec
    q.put((rank, grec.numpy(), gcnt.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gather_equals_single_process():
    _run_world(7, 2)                           # ragged shards: 4 + 3 (padded collective)
    _run_world(6, 2)                           # even shards: the gathered buffer is returned as is


def _run_world(n_items, world):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    os.environ["PYTHONPATH"] = root + os.pathsep + os.environ.get("PYTHONPATH", "")
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_items, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ref_rec, ref_cnt = _fake_detect(list(range(n_items)))
    for _, rec, cnt in results:
        assert np.array_equal(rec, ref_rec.numpy()) and np.array_equal(cnt, ref_cnt.numpy())


def test_single_process_passthrough():
    from livecell_instance_segmentation_b200.dist import all_gather_detections
    rec, cnt = _fake_detect([0, 1, 2])
    r2, c2 = all_gather_detections(rec, cnt, 3)
    assert r2 is rec and c2 is cnt
    r3, c3 = all_gather_detections(rec, cnt, 3, async_op=True).wait()
    assert r3 is rec and c3 is cnt

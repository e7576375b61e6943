"""Multi-GPU sharding of the region pipeline.

The path is embarrassingly parallel over images (src/custom_maskrcnn.py:164: the loop body has no
cross-image dependence), so images are split into contiguous per-rank blocks with NO data-path
collective; the only exchange is one all-gather of the fixed-size detection records
[imgs_per_rank, D, 6] (+ counts) at the end.  One process per GPU (torchrun); NCCL on GPUs, gloo on
CPU for the host-logic tests.
"""
from __future__ import annotations

from typing import Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of rank `rank`; blocks differ in size by at most one item."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_shard(n_items: int, world_size: int) -> int:
    return (n_items + world_size - 1) // world_size


def all_gather_detections(records: torch.Tensor, counts: torch.Tensor, n_items: int, group=None):
    """records [n_local, D, 6] f32, counts [n_local] i32 of this rank's shard  ->
    (records [n_items, D, 6], counts [n_items]) identical on every rank, in global image order.
    Shards are padded to the largest shard so the collective has a fixed size."""
    if not dist.is_initialized():
        return records, counts
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    cap = max_shard(n_items, world)
    n_local, D = records.shape[0], records.shape[1]
    rec = records.new_zeros((cap, D, 6))
    cnt = counts.new_zeros((cap,))
    rec[:n_local] = records
    cnt[:n_local] = counts
    out_rec = records.new_empty((world * cap, D, 6))
    out_cnt = counts.new_empty((world * cap,))
    dist.all_gather_into_tensor(out_rec, rec, group=group)
    dist.all_gather_into_tensor(out_cnt, cnt, group=group)
    recs, cnts = [], []
    for r in range(world):
        s, e = shard_range(n_items, r, world)
        recs.append(out_rec[r * cap: r * cap + (e - s)])
        cnts.append(out_cnt[r * cap: r * cap + (e - s)])
    return torch.cat(recs, dim=0), torch.cat(cnts, dim=0)

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


_last_bind_report: dict = {}


def last_bind_report() -> dict:
    """Why the last bind_to_gpu_numa_node() did what it did (for bench.py's e2e line)."""
    return dict(_last_bind_report)


def bind_to_gpu_numa_node(device_index: int) -> list:
    """Pin this process to the CPU cores NVML reports as local to GPU `device_index` (restricted to the cores the
    process may use), so that pinned host buffers are first-touched on the GPU's own NUMA node and the host->device
    copies of 8 ranks do not cross the socket interconnect.  Returns the core list ([] when nothing was changed);
    last_bind_report() says why."""
    import os
    rep = {"device_index": device_index, "bound": 0}
    _last_bind_report.clear()
    _last_bind_report.update(rep)
    try:
        allowed = os.sched_getaffinity(0)
        _last_bind_report["allowed_cores"] = len(allowed)
        try:
            nodes = [d for d in os.listdir("/sys/devices/system/node") if d.startswith("node") and d[4:].isdigit()]
            _last_bind_report["numa_nodes"] = len(nodes)
        except OSError:
            _last_bind_report["numa_nodes"] = None
        import pynvml
        pynvml.nvmlInit()
        visible = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = device_index
        if visible:
            ids = [v.strip() for v in visible.split(",") if v.strip()]
            if device_index < len(ids) and ids[device_index].isdigit():
                idx = int(ids[device_index])
        handle = pynvml.nvmlDeviceGetHandleByIndex(idx)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(handle, (ncpu + 63) // 64)
        local = {i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1}
        cores = sorted(local & allowed)
        _last_bind_report["gpu_local_cores"] = len(local)
        _last_bind_report["gpu_local_and_allowed"] = len(cores)
        if not cores:
            _last_bind_report["reason"] = "none of the GPU's local cores is in this process's affinity mask"
        elif len(cores) >= len(allowed):
            _last_bind_report["reason"] = ("NVML reports every allowed core as local to this GPU "
                                           "(single NUMA node / VM without topology): nothing to narrow")
        else:
            os.sched_setaffinity(0, cores)
            _last_bind_report["bound"] = len(cores)
            _last_bind_report["reason"] = "bound to the GPU's local cores"
            return cores
    except Exception as exc:  # NVML missing, no permission, ...
        _last_bind_report["reason"] = f"not attempted: {type(exc).__name__}: {exc}"
    return []


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous block [start, end) of rank `rank`; blocks differ in size by at most one item."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("bad rank/world_size")
    base, rem = divmod(n_items, world_size)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def max_shard(n_items: int, world_size: int) -> int:
    return (n_items + world_size - 1) // world_size


class GatherHandle:
    """Pending all-gather of detection records; wait() makes the current stream wait and returns the tensors."""

    def __init__(self, works, finish):
        self._works, self._finish, self._result = works, finish, None

    def wait(self):
        if self._result is None:
            for w in self._works:
                w.wait()
            self._result = self._finish()
        return self._result


def all_gather_detections(records: torch.Tensor, counts: torch.Tensor, n_items: int, group=None, async_op: bool = False):
    """records [n_local, D, 6] f32, counts [n_local] i32 of this rank's shard  ->
    (records [n_items, D, 6], counts [n_items]) identical on every rank, in global image order.
    Shards are padded to the largest shard so the collective has a fixed size.  With async_op the two
    collectives are only enqueued (NCCL's own stream): the caller keeps launching the next batch and calls
    .wait() on the returned GatherHandle when it needs the result — the records are 12 KB/frame, so the
    exchange then costs nothing on the critical path."""
    if not dist.is_initialized():
        return GatherHandle([], lambda: (records, counts)) if async_op else (records, counts)
    world = dist.get_world_size(group)
    cap = max_shard(n_items, world)
    n_local, D = records.shape[0], records.shape[1]
    even = n_items == world * cap and n_local == cap
    if even:
        rec, cnt = records.contiguous(), counts.contiguous()
    else:
        rec = records.new_zeros((cap, D, 6))
        cnt = counts.new_zeros((cap,))
        rec[:n_local] = records
        cnt[:n_local] = counts
    out_rec = records.new_empty((world * cap, D, 6))
    out_cnt = counts.new_empty((world * cap,))
    works = [dist.all_gather_into_tensor(out_rec, rec, group=group, async_op=async_op),
             dist.all_gather_into_tensor(out_cnt, cnt, group=group, async_op=async_op)]

    def finish():
        if even:
            return out_rec, out_cnt
        recs, cnts = [], []
        for r in range(world):
            s, e = shard_range(n_items, r, world)
            recs.append(out_rec[r * cap: r * cap + (e - s)])
            cnts.append(out_cnt[r * cap: r * cap + (e - s)])
        return torch.cat(recs, dim=0), torch.cat(cnts, dim=0)

    if async_op:
        return GatherHandle(works, finish)
    return finish()

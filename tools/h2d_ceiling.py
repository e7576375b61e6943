#!/usr/bin/env python
"""What does the host->device path of this box give N concurrent ranks?  

    python tools/h2d_ceiling.py                                                   # one GPU
    python -m torch.distributed.run --nproc-per-node N ... tools/h2d_ceiling.py   # N ranks at once

Every rank copies the e2e step's input volume (1.65 GB: the fp32 level-0 feature maps dominate) from its own pinned buffers
with plain Tensor.copy_ (one cudaMemcpyAsync per call; no batching API), nothing else running on the GPUs, for a few
configurations (1-4 copy streams, 8/64-frame chunks).  Rank 0 prints one JSON line per configuration with the per-rank
rates and the aggregate: this is the ceiling bench.py's `e2e.h2d_GBps` is quoted against."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import numpy as np
    import torch
    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    F, FH, FW, C = 64, 130, 176, 256
    host = torch.empty((F, FH, FW, C), dtype=torch.float32).pin_memory()
    host.normal_()
    dst = torch.empty((F, FH, FW, C), dtype=torch.float32, device=dev)
    nbytes = host.numel() * 4
    for streams, chunk in ((1, 64), (1, 8), (2, 8), (4, 8), (4, 2)):
        ss = [torch.cuda.Stream() for _ in range(streams)]

        def copy_all():
            for i, f0 in enumerate(range(0, F, chunk)):
                with torch.cuda.stream(ss[i % streams]):
                    dst[f0:f0 + chunk].copy_(host[f0:f0 + chunk], non_blocking=True)
            for s in ss:
                torch.cuda.current_stream().wait_stream(s)
        copy_all()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        ms = []
        for _ in range(5):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            e0.record()
            copy_all()
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        rate = nbytes / 1e9 / (float(np.median(ms)) * 1e-3)
        rates = [rate]
        if world > 1:
            t = torch.tensor([rate], device=dev, dtype=torch.float64)
            out = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(out, t)
            rates = [float(o.item()) for o in out]
        if rank == 0:
            print(json.dumps({"n_ranks": world, "copy_streams": streams, "chunk_frames": chunk, "bytes_per_rank": nbytes,
                              "GBps_per_rank": rates, "GBps_min": min(rates), "GBps_aggregate": sum(rates),
                              "host_cores": len(os.sched_getaffinity(0))}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

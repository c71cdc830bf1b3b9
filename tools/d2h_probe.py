#!/usr/bin/env python
"""Device→host copy bandwidth per rank while all ranks copy at once (development aid: explains how bench.py's `e2e`
scales with the number of GPUs — every step returns 192 MB of node series per GPU to pinned host memory).

  python tools/d2h_probe.py                                                    one GPU
  python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/d2h_probe.py   N GPUs, concurrently
"""
import json
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import pin_to_gpu_numa_node  # noqa: E402


def main() -> None:
    rank, world, local = (int(os.environ.get(k, d)) for k, d in (("RANK", 0), ("WORLD_SIZE", 1), ("LOCAL_RANK", 0)))
    if world > 1:
        dist.init_process_group("gloo", rank=rank, world_size=world)
        if "--no-pin" not in sys.argv:
            pin_to_gpu_numa_node(local)
    torch.cuda.set_device(local)
    n = 192_000_000 // 8
    dev = torch.zeros(n, dtype=torch.float64, device="cuda")
    host = torch.empty(n, dtype=torch.float64).pin_memory()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = {}
    for name, (dst, src) in {"d2h": (host, dev), "h2d": (dev, host)}.items():
        for _ in range(3):
            dst.copy_(src, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0.record()
        for _ in range(20):
            dst.copy_(src, non_blocking=True)
        e1.record()
        torch.cuda.synchronize()
        out[name + "_GBps"] = round(20 * n * 8 / (e0.elapsed_time(e1) * 1e-3) / 1e9, 2)
    out.update(rank=rank, world=world, cpus=len(os.sched_getaffinity(0)))
    rows = [out]
    if world > 1:
        rows = [None] * world
        dist.all_gather_object(rows, out)
    if rank == 0:
        for r in rows:
            print(json.dumps(r))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

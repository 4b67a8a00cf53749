#!/usr/bin/env python
"""Aggregate host->device bandwidth of one box when N ranks copy at the same time (pinned memory, one GPU per rank).
Launched under torchrun; rank 0 prints one JSON line.  Diagnostic for the end-to-end scaling of the host entry points:
the per-GPU H2D rate at N=8 bounds the e2e throughput of a path that uploads ~20 KB per window."""
import json
import os
import time

import torch
import torch.distributed as dist


def main():
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    lr = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(lr)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", lr))
    nbytes = 256 << 20
    h = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
    h.fill_(7)
    d = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
    h2 = torch.empty(16 << 20, dtype=torch.uint8).pin_memory()
    d2 = torch.zeros(16 << 20, dtype=torch.uint8, device="cuda")
    st = torch.cuda.Stream()
    st2 = torch.cuda.Stream()
    res = {}
    for name, both in (("h2d", False), ("h2d+d2h", True)):
        for _ in range(3):
            d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 20
        with torch.cuda.stream(st):
            for _ in range(reps):
                d.copy_(h, non_blocking=True)
        if both:
            with torch.cuda.stream(st2):
                for _ in range(reps):
                    h2.copy_(d2, non_blocking=True)
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        gbs = torch.tensor([reps * nbytes / dt / 1e9], dtype=torch.float64, device="cuda")
        mn = gbs.clone()
        if world > 1:
            dist.all_reduce(gbs, op=dist.ReduceOp.SUM)
            dist.all_reduce(mn, op=dist.ReduceOp.MIN)
        res[name] = {"aggregate_gbs": round(float(gbs), 1), "min_rank_gbs": round(float(mn), 1)}
    if rank == 0:
        print(json.dumps({"n": world, "cpus": os.cpu_count(), "affinity": len(os.sched_getaffinity(0)),
                          "omp": os.environ.get("OMP_NUM_THREADS"), **res}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

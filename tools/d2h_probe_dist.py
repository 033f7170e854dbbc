#!/usr/bin/env python
"""Concurrent device->host bandwidth of all ranks of one box (what bounds the end-to-end leg of the
benchmark at N > 1: every site record crosses PCIe into one host's memory).  Under torchrun; each
rank copies a device buffer into its own pinned host buffer, all ranks at once, then alone.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 \
        --master-port 29541 tools/d2h_probe_dist.py [--gb 2]
"""
import argparse
import json
import os
import time


def main():
    import torch
    import torch.distributed as dist
    ap = argparse.ArgumentParser()
    ap.add_argument("--gb", type=float, default=2.0)
    a = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", 0)))
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = int(a.gb * (1 << 30))
    d = torch.empty(n, dtype=torch.uint8, device=dev)
    h = torch.empty(n, dtype=torch.uint8).pin_memory()

    def run(reps, direction):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            if direction == "d2h":
                h.copy_(d, non_blocking=True)
            else:
                d.copy_(h, non_blocking=True)
        torch.cuda.synchronize()
        return reps * n / (time.perf_counter() - t0) / 1e9

    out = {}
    for direction in ("d2h", "h2d"):
        run(1, direction)
        together = run(4, direction)
        t = torch.tensor([together], dtype=torch.float64, device=dev)
        alone = torch.zeros(world, dtype=torch.float64, device=dev)
        for r in range(world):              # one rank at a time
            if world > 1:
                dist.barrier()
            if r == rank:
                alone[r] = run(2, direction) if world == 1 else _solo(h, d, n, direction)
        if world > 1:
            g = [torch.zeros_like(t) for _ in range(world)]
            dist.all_gather(g, t)
            dist.all_reduce(alone)
            per = [float(x.item()) for x in g]
        else:
            per = [together]
        out[direction] = {"concurrent_gb_per_s_by_rank": [round(x, 1) for x in per],
                          "concurrent_aggregate_gb_per_s": round(sum(per), 1),
                          "alone_gb_per_s_by_rank": [round(float(x), 1) for x in alone.tolist()]}
    if rank == 0:
        out["n_gpus"] = world
        out["buffer_gb"] = a.gb
        out["cpus"] = os.cpu_count()
        print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


def _solo(h, d, n, direction):
    import torch
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(2):
        if direction == "d2h":
            h.copy_(d, non_blocking=True)
        else:
            d.copy_(h, non_blocking=True)
    torch.cuda.synchronize()
    return 2 * n / (time.perf_counter() - t0) / 1e9


if __name__ == "__main__":
    main()

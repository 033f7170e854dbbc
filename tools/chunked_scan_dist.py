#!/usr/bin/env python
"""BASELINE configs[3] at full per-GPU size: one long sequence cut into contiguous chunks with an
(m - 1)-base halo, one chunk per rank (1.25 Gbp each by default = 10 Gbp over 8 GPUs), each rank
generates its chunk on the device, scans it (background record + hit list) and ONE all-reduce
sums the background record.  Launch under torchrun; rank 0 prints a JSON line.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 \
        --master-port 29530 tools/chunked_scan_dist.py [--bp-per-rank 1250000000]
"""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch
    import torch.distributed as dist

    import bench
    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, _cabi
    from cgb_b200 import dist as cdist
    from cgb_b200 import motif_batch as mb

    ap = argparse.ArgumentParser()
    ap.add_argument("--bp-per-rank", type=int, default=1_250_000_000)
    ap.add_argument("--reps", type=int, default=3)
    ap.add_argument("--two-scans", action="store_true",
                    help="round-1 flow: a stats scan, the all-reduce, then a separate filter scan + exact re-score")
    args = ap.parse_args()
    rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    d, site_lists, weights = bench.lexa()
    M = PSSMModel([SiteCollection(s) for s in site_lists], weights, device=dev)
    m = M.length
    total = args.bp_per_rank * world
    starts, ends, counts = cdist.chunk_bounds(total, m, world)
    n = int(ends[rank] - starts[rank])
    # the chunk's bases: a counter-based generator keyed by the global position block, so that
    # the halo of rank r equals the head of rank r + 1
    g = PackedGenome([n], dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    ascii_all = torch.empty(n, dtype=torch.uint8, device=dev)
    block = 1 << 26
    for b0 in range(int(starts[rank]) // block * block, int(ends[rank]), block):
        gen = torch.Generator(device=dev)
        gen.manual_seed(1000 + b0 // block)
        vals = lut[torch.randint(0, 4, (block,), device=dev, generator=gen)]
        lo, hi = max(b0, int(starts[rank])), min(b0 + block, int(ends[rank]))
        ascii_all[lo - int(starts[rank]):hi - int(starts[rank])] = vals[lo - b0:hi - b0]
    g.pack_into(0, ascii_all)
    torch.cuda.synchronize()
    del ascii_all
    thr = M.threshold()
    times = []
    for rep in range(args.reps + 1):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        if args.two_scans:
            bg = M.background_stats(g)
            if world > 1:
                cdist.allreduce_background(bg)
            pending = M._launch_scan(g, thr)
            b.record()
            torch.cuda.synchronize()
            if rep:
                times.append(a.elapsed_time(b))
            seq, pos, strand, score = M._collect_scan(g, pending)
            n_sites = len(pos)
        else:
            # ONE fused pass: background record + candidate flags, then the ordered site list; the
            # record of the chunk goes through the one all-reduce
            res = mb.scan_batch([M], g, [thr], n_bins=256, to_host=False)
            if world > 1:
                cdist.allreduce_batch(res["stats"], res["hist"])
            b.record()
            torch.cuda.synchronize()
            if rep:
                times.append(a.elapsed_time(b))
            bg = {"stats": res["stats"][0]}
            n_sites = int(res["hits"][0].numel())
    t = torch.tensor([float(np.mean(times))], dtype=torch.float64, device=dev)
    nh = torch.tensor([n_sites], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(nh, op=dist.ReduceOp.SUM)
    stats = bg["stats"].cpu().numpy()
    if rank == 0:
        ms = float(t.item())
        nwin = total - m + 1
        print(json.dumps({
            "workload": "configs[3]: one %d bp sequence, %d chunks with a %d-base halo, LexA 22-bp PSWM: "
                        "background record (one all-reduce) + hit list" % (total, world, m - 1),
            "n_gpus": world, "ms_max_over_ranks": ms, "windows": nwin,
            "bp_motif_per_s": (2 if args.two_scans else 1) * nwin / (ms * 1e-3),
            "bp_motif_per_s_note": "both scans counted: windows * 2 / time" if args.two_scans else
                                   "one fused pass per window (background record + site candidates): windows / time",
            "background": {"n": float(stats[_cabi.BG_N]), "mean": float(stats[_cabi.BG_MEAN]),
                           "std": float(stats[_cabi.BG_STD])},
            "hits_total": int(nh.item())}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()

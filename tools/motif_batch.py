#!/usr/bin/env python
"""BASELINE configs[4] style measurement: many JASPAR-scale motifs (8-30 bp) over one
synthetic sequence, all hit scans enqueued back to back (cgb_b200.scan_hits_many).

    python tools/motif_batch.py [--motifs 100] [--bp 100000000]

Reports aggregate bp*motif/s (CUDA events around the whole batch, host collection excluded
and included) and the algorithmic lookup-add rate against the issue roofline.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, scan_hits_many

    ap = argparse.ArgumentParser()
    ap.add_argument("--motifs", type=int, default=100)
    ap.add_argument("--bp", type=int, default=100_000_000)
    ap.add_argument("--many", action="store_true",
                    help="go through scan_hits_many (memory-budgeted launch/collect) instead of timing the bare launches")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    rng = np.random.default_rng(5)
    models = []
    for k in range(args.motifs):
        m = int(rng.integers(8, 31))
        cols = rng.dirichlet([0.5] * 4, size=m)
        sites = ["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m)) for _ in range(20)]
        models.append(PSSMModel([SiteCollection(sites)], [1.0], device=dev))
    t0 = time.perf_counter()
    thresholds = [M.threshold() for M in models]             # Patser DP, host C++
    t_thr = time.perf_counter() - t0
    gen = torch.Generator(device=dev)
    gen.manual_seed(5)
    g = PackedGenome([args.bp], dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    g.pack_into(0, lut[torch.randint(0, 4, (args.bp,), device=dev, generator=gen)])
    torch.cuda.synchronize()

    scan_hits_many(models[:4], g, thresholds[:4])            # warm up (table uploads, allocator)
    torch.cuda.synchronize()
    if args.many:
        t0 = time.perf_counter()
        res = scan_hits_many(models, g, thresholds)
        wall = time.perf_counter() - t0
        units = sum(g.total_windows(M.length) for M in models)
        print(json.dumps({"motifs": args.motifs, "bp": args.bp, "via": "scan_hits_many", "wall_s": wall,
                          "bp_motif_per_s_wall": units / wall, "hits": int(sum(len(r[1]) for r in res)),
                          "peak_mem_gb": torch.cuda.max_memory_allocated(dev) / 1e9}))
        return
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    a.record()
    pending = [M._launch_scan(g, thr) for M, thr in zip(models, thresholds)]
    b.record()
    torch.cuda.synchronize()
    gpu_ms = a.elapsed_time(b)
    res = [M._collect_scan(g, p) for M, p in zip(models, pending)]
    wall = time.perf_counter() - t0
    units = sum(g.total_windows(M.length) for M in models)
    adds = sum(2 * M.length * g.total_windows(M.length) for M in models)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    print(json.dumps({
        "motifs": args.motifs, "bp": args.bp, "motif_lengths": "8-30",
        "gpu_ms": gpu_ms, "bp_motif_per_s_gpu": units / (gpu_ms * 1e-3),
        "issue_frac": adds / (gpu_ms * 1e-3) / (sms * 128 * 1965e6),
        "wall_s_incl_collect": wall, "bp_motif_per_s_wall": units / wall,
        "hits": int(sum(len(r[1]) for r in res)), "patser_thresholds_s": t_thr}))


if __name__ == "__main__":
    main()

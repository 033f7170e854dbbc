#!/usr/bin/env python
"""Per-motif cost of the fused scan's site pass (candidate re-score + ordered emission) against
the candidate density: motifs of the benchmark set picked across the density range, each run alone
through cgbk_scan_batch; step time (CUDA events) minus the scan kernel's own time.

    python tools/site_pass_times.py [--bp 1000000000]
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

    import bench
    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, _cabi, prepare_thresholds
    from cgb_b200 import motif_batch as mb
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=1_000_000_000)
    ap.add_argument("--pick", type=int, default=14)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    sl = bench.motif_site_lists(500)
    Ms = [PSSMModel([SiteCollection(s)], [1.0], device=dev) for s in sl]
    prepare_thresholds(Ms)
    e = np.array([mb.expected_hits(M, a.bp) for M in Ms])
    order = np.argsort(e)
    picks = [order[int(round(x))] for x in np.linspace(0, len(order) - 1, a.pick)]
    g = PackedGenome([a.bp], dev)
    g.pack_into(0, bench.generate_bases(0, a.bp, dev))
    torch.cuda.synchronize()
    _cabi.profile_enable(2)
    rows = []
    for k in picks:
        M = Ms[k]
        batch = mb.ModelBatch([M], dev)
        cap = int(e[k] * 1.3 + 1e6)
        acc = np.zeros(1)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        reps = 4
        for it in range(2 + reps):
            if it == 2:
                acc[:] = 0
                ev[0].record()
            res = mb.scan_batch([M], g, [M.threshold()], n_bins=256, to_host=False, recycle=True, hit_cap=cap,
                                kernel_ms=acc)
        ev[1].record()
        torch.cuda.synchronize()
        step = ev[0].elapsed_time(ev[1]) / reps
        kern = float(acc[0]) / reps
        rows.append({"m": M.length, "hits": int(res["hits"][0]), "cands_per_tile": round(float(e[k]) / a.bp * 4064, 2),
                     "scan_ms": round(kern, 3), "site_pass_ms": round(step - kern, 3)})
        print(json.dumps(rows[-1]))
        del batch
        mb.release_buffers(dev)


if __name__ == "__main__":
    main()

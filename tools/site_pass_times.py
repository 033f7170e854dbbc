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
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    sl = bench.motif_site_lists(500)
    Ms = [PSSMModel([SiteCollection(s)], [1.0], device=dev) for s in sl]
    prepare_thresholds(Ms)
    e = np.array([mb.expected_hits(M, a.bp) for M in Ms])
    g = PackedGenome([a.bp], dev)
    g.pack_into(0, bench.generate_bases(0, a.bp, dev))
    torch.cuda.synchronize()
    batch = mb.ModelBatch(Ms, dev)
    thr = [M.threshold() for M in Ms]
    cap = int(min(e.sum() * 1.3, 1.5e9))
    _cabi.profile_enable(2)
    kern, site = np.zeros(len(Ms)), np.zeros(len(Ms))
    reps = 2
    for it in range(1 + reps):
        if it == 1:
            kern[:] = 0
            site[:] = 0
        res = mb.scan_batch(Ms, g, thr, n_bins=256, to_host=False, recycle=True, hit_cap=cap, kernel_ms=kern,
                            site_ms=site)
    kern /= reps
    site /= reps
    hits = np.array(res["hits"], dtype=np.float64)
    dens = hits / a.bp * 4064
    print(json.dumps({"motifs": len(Ms), "scan_ms_total": round(float(kern.sum()), 1),
                      "site_pass_ms_total": round(float(site.sum()), 1), "hits": float(hits.sum())}))
    edges = [0, 1, 4, 16, 32, 64, 256, 1024, 4097]
    for lo, hi in zip(edges[:-1], edges[1:]):
        sel = (dens >= lo) & (dens < hi)
        if sel.any():
            print(json.dumps({"sites_per_tile": "[%d, %d)" % (lo, hi), "motifs": int(sel.sum()),
                              "hits": float(hits[sel].sum()), "site_pass_ms_sum": round(float(site[sel].sum()), 2),
                              "site_pass_ms_mean": round(float(site[sel].mean()), 4),
                              "ns_per_hit": round(float(site[sel].sum() * 1e6 / max(hits[sel].sum(), 1)), 4)}))


if __name__ == "__main__":
    main()

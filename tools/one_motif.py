#!/usr/bin/env python
"""Run the fused scan + site pass of ONE motif of the benchmark set a few times (a target for ncu):
--rank r picks the motif with the r-th most expected sites (0 = densest)."""
import argparse
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, prepare_thresholds
    from cgb_b200 import motif_batch as mb
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=1_000_000_000)
    ap.add_argument("--rank", type=int, default=20)
    ap.add_argument("--reps", type=int, default=3)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    sl = bench.motif_site_lists(500)
    Ms = [PSSMModel([SiteCollection(s)], [1.0], device=dev) for s in sl]
    prepare_thresholds(Ms)
    e = np.array([mb.expected_hits(M, a.bp) for M in Ms])
    k = int(np.argsort(-e)[a.rank])
    M = Ms[k]
    print("motif", k, "m", M.length, "expected sites %.3g" % e[k])
    g = PackedGenome([a.bp], dev)
    g.pack_into(0, bench.generate_bases(0, a.bp, dev))
    batch = mb.ModelBatch([M], dev)
    for _ in range(a.reps):
        res = mb.scan_batch([M], g, [M.threshold()], n_bins=256, to_host=False, recycle=True,
                            hit_cap=int(e[k] * 1.3 + 1e6))
    torch.cuda.synchronize()
    print("sites", res["hits"][0])


if __name__ == "__main__":
    main()

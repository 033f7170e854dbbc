#!/usr/bin/env python
"""Host-side cost of the per-model calls a cgb run makes once per genome (4 Mbp genome): model
construction, score_self, threshold, hit scan (launch / collect), one-call pipeline."""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import bench
    from cgb_b200 import GeneTable, PackedGenome, PSSMModel, SiteCollection

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    d, site_lists, weights = bench.lexa()
    colls = [SiteCollection(s) for s in site_lists]
    bp = 4_000_000
    pinned = torch.from_numpy(bench.synthetic_genome(bp, 3)).pin_memory()
    gs, ge, gstrand = bench.gene_layout(bp, 2400)
    ps, pe = GeneTable(gs, ge, gstrand, bp).promoter_region_location(300, 50)
    g = PackedGenome.from_sequences([None], dev, pinned=[pinned])
    acc = {}

    def tick(name, t0):
        torch.cuda.synchronize()
        acc.setdefault(name, []).append((time.perf_counter() - t0) * 1e3)

    rng = np.random.default_rng(0)
    for it in range(12):
        w = rng.dirichlet([1.0] * len(colls)).tolist()
        t0 = time.perf_counter(); M = PSSMModel(colls, w, device=dev); M._model(); tick("PSSMModel() + device tables", t0)
        t0 = time.perf_counter(); M.threshold(); tick("threshold (Patser DP)", t0)
        t0 = time.perf_counter(); M.score_self(); tick("score_self", t0)
        t0 = time.perf_counter(); M.regulation_pipeline(pinned, ps, pe, 0.03, d["alpha"]); tick("regulation_pipeline (1st)", t0)
        t0 = time.perf_counter(); M.regulation_pipeline(pinned, ps, pe, 0.03, d["alpha"]); tick("regulation_pipeline (2nd)", t0)
        t0 = time.perf_counter(); p = M._launch_scan(g); tick("scan launch", t0)
        t0 = time.perf_counter(); M._collect_scan(g, p); tick("scan collect", t0)
        t0 = time.perf_counter(); M.scan_hits(g); tick("scan_hits (2nd)", t0)
    print(json.dumps({k: round(float(np.median(v[2:])), 3) for k, v in acc.items()}, indent=1))


if __name__ == "__main__":
    main()

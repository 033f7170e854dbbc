#!/usr/bin/env python
"""Host-side phase timing of the config-2 end-to-end step (where do the microseconds go)."""
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
    from cgb_b200 import GeneTable, PackedGenome, PSSMModel, SiteCollection, _cabi

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    d, site_lists, weights = bench.lexa()
    prior, alpha = d["prior_regulation_probability"], d["alpha"]
    M = PSSMModel([SiteCollection(s) for s in site_lists], weights, device=dev)
    host = bench.synthetic_genome(bench.GENOME_BP, 2)
    pinned = torch.from_numpy(host).pin_memory()
    gs, ge, gstrand = bench.gene_layout(bench.GENOME_BP, bench.N_GENES)
    gt = GeneTable(gs, ge, gstrand, bench.GENOME_BP)
    ps, pe = gt.promoter_region_location(bench.UP, bench.DOWN)
    post = torch.empty(bench.N_GENES, dtype=torch.float64, device=dev)
    post_host = torch.empty(bench.N_GENES, dtype=torch.float64).pin_memory()

    def sync():
        torch.cuda.synchronize()
        return time.perf_counter()

    acc = {}
    for it in range(25):
        t = [sync()]
        gobj = PackedGenome(np.array([pinned.numel()]), dev); t.append(time.perf_counter())
        devb = pinned.to(dev, non_blocking=True); t.append(time.perf_counter())
        gobj.pack_into(0, devb); t.append(time.perf_counter())
        t.append(sync())
        b = M.background_stats(gobj, n_bins=256, keep_scores=True); t.append(time.perf_counter())
        M.fit_background(bg=b, sync=False); t.append(time.perf_counter())
        t.append(sync())
        desc = M.promoter_descriptors(gobj, ps, pe); t.append(time.perf_counter())
        p = M.binding_probabilities(gobj, None, None, prior, alpha, out=post, region_tensors=desc)
        t.append(time.perf_counter())
        t.append(sync())
        post_host.copy_(p, non_blocking=True); torch.cuda.current_stream().synchronize(); t.append(time.perf_counter())
        names = ["PackedGenome()", "h2d issue", "pack issue", "h2d+pack gpu wait", "bg_stats issue",
                 "fit issue", "bg gpu wait", "descriptors", "posterior issue", "posterior gpu wait", "d2h"]
        if it >= 5:
            for n, a, b2 in zip(names, t[:-1], t[1:]):
                acc.setdefault(n, []).append((b2 - a) * 1e6)
    for n, v in acc.items():
        print("%-22s %8.1f us" % (n, float(np.mean(v))))
    print("%-22s %8.1f us" % ("sum", sum(float(np.mean(v)) for v in acc.values())))


if __name__ == "__main__":
    main()

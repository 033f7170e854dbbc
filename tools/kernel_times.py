#!/usr/bin/env python
"""Device time of each stage of the config-2 step (stats scan call, posterior call), warm L2, replayed
from a CUDA graph so that host overhead does not count.  For A/B-ing kernel variants (CGBK_LIB=...)."""
import json
import os
import sys

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
    prior, alpha = d["prior_regulation_probability"], d["alpha"]
    M = PSSMModel([SiteCollection(s) for s in site_lists], weights, device=dev)
    host = bench.synthetic_genome(bench.GENOME_BP, 2)
    g = PackedGenome.from_sequences([host], dev)
    gs, ge, gstrand = bench.gene_layout(bench.GENOME_BP, bench.N_GENES)
    gt = GeneTable(gs, ge, gstrand, bench.GENOME_BP)
    ps, pe = gt.promoter_region_location(bench.UP, bench.DOWN)
    desc = M.promoter_descriptors(g, ps, pe)
    bg = M.background_stats(g, n_bins=256, keep_scores=True)
    M.fit_background(bg=bg)
    post = torch.empty(bench.N_GENES, dtype=torch.float64, device=dev)
    reps = 200

    def timeit(fn, per_graph=20):
        """``per_graph`` calls captured into one CUDA graph and replayed: the host (Python +
        driver, ~12 us per call) is out of the picture, what is left is device time per call."""
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        side = torch.cuda.Stream()
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.stream(side):
            fn()
            side.synchronize()
            with torch.cuda.graph(graph, stream=side):
                for _ in range(per_graph):
                    fn()
            graph.replay()
            side.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(side)
            for _ in range(reps // per_graph):
                graph.replay()
            b.record(side)
            side.synchronize()
        return a.elapsed_time(b) / (reps // per_graph * per_graph) * 1e3

    out = {
        "stats_call_us": timeit(lambda: M.background_stats(g, out=bg, keep_scores=True)),
        "posterior_plane_us": timeit(lambda: M.binding_probabilities(g, None, None, prior, alpha, out=post,
                                                                     region_tensors=desc)),
        "posterior_rescore_us": timeit(lambda: M.binding_probabilities(g, None, None, prior, alpha, out=post,
                                                                       region_tensors=desc, use_plane=False)),
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()

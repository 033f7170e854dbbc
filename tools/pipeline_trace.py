#!/usr/bin/env python
"""Phase timestamps of cgbk_regulation_pipeline on the config-2 inputs: builds the library
variant with -DCGBK_PIPE_TRACE into a scratch path and runs a few calls (stderr shows, per phase,
when the device reached it and when the host enqueued it)."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    # the library path is read when cgb_b200 is first imported
    os.environ["CGBK_LIB"] = os.path.join(ROOT, "cgb_b200", "_lib", "libcgbk_trace.so")
    from cgb_b200 import build
    build.build_library(defines=("-DCGBK_PIPE_TRACE",), out="libcgbk_trace.so")
    if "--build-only" in sys.argv:
        return
    import torch

    import bench
    from cgb_b200 import GeneTable, PSSMModel, SiteCollection

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    d, site_lists, weights = bench.lexa()
    prior, alpha = d["prior_regulation_probability"], d["alpha"]
    M = PSSMModel([SiteCollection(s) for s in site_lists], weights, device=dev)
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    bp = int(args[0]) if args else bench.GENOME_BP
    host = bench.synthetic_genome(bp, 2)
    pinned = torch.from_numpy(host).pin_memory()
    gs, ge, gstrand = bench.gene_layout(bp, bench.N_GENES)
    gt = GeneTable(gs, ge, gstrand, bp)
    ps, pe = gt.promoter_region_location(bench.UP, bench.DOWN)
    src = pinned
    if "--packed" in sys.argv:
        from cgb_b200 import HostPackedGenome
        src = HostPackedGenome.from_ascii(pinned)
    import time
    for i in range(6):
        sys.stderr.write("--- call %d\n" % i)
        t0 = time.perf_counter()
        M.regulation_pipeline(src, ps, pe, prior, alpha)
        sys.stderr.write("python call %.1f us\n" % ((time.perf_counter() - t0) * 1e6))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""End-to-end step time of cgbk_regulation_pipeline on the config-2 inputs (pinned host buffers,
wall clock per call, median of many).  CGBK_PIPE_CHUNKS=1|3|4 selects the upload chunking."""
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
    from cgb_b200 import GeneTable, PSSMModel, SiteCollection

    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    d, site_lists, weights = bench.lexa()
    prior, alpha = d["prior_regulation_probability"], d["alpha"]
    M = PSSMModel([SiteCollection(s) for s in site_lists], weights, device=dev)
    bp = int(sys.argv[1]) if len(sys.argv) > 1 else bench.GENOME_BP
    pinned = torch.from_numpy(bench.synthetic_genome(bp, 2)).pin_memory()
    gs, ge, gstrand = bench.gene_layout(bp, bench.N_GENES)
    ps, pe = GeneTable(gs, ge, gstrand, bp).promoter_region_location(bench.UP, bench.DOWN)
    devbuf = torch.empty(bp, dtype=torch.uint8, device=dev)
    out = {"chunks": os.environ.get("CGBK_PIPE_CHUNKS", "auto"), "bp": bp}
    for _ in range(20):
        M.regulation_pipeline(pinned, ps, pe, prior, alpha)
    t = []
    for _ in range(200):
        t0 = time.perf_counter()
        M.regulation_pipeline(pinned, ps, pe, prior, alpha)
        t.append((time.perf_counter() - t0) * 1e6)
    out["pipeline_us_median"] = float(np.median(t))
    out["pipeline_us_p10"] = float(np.percentile(t, 10))
    # the C call alone (arguments marshalled once): what the Python wrapper adds on top
    from cgb_b200 import _cabi
    lib = _cabi.lib()
    starts = np.ascontiguousarray(ps, dtype=np.int64)
    ends = np.ascontiguousarray(pe, dtype=np.int64)
    n = len(starts)
    scratch, host = M._pipe_scratch, M._pipe_host
    stage_bytes = int(lib.cgbk_pipeline_staging_bytes(n))
    hp = host.data_ptr()
    lo, hi = M._score_bounds
    args = (M._model(), pinned.data_ptr(), bp, starts.ctypes.data, ends.ctypes.data, n, float(M._mu_m),
            float(M._std_m), float(prior), float(alpha), float(lo), float(hi), 256, scratch.data_ptr(),
            scratch.numel(), hp, stage_bytes, hp + stage_bytes, hp + stage_bytes + 8 * n,
            hp + stage_bytes + 8 * n + 64, M._stream())
    t = []
    for _ in range(200):
        t0 = time.perf_counter()
        rc = lib.cgbk_regulation_pipeline(*args)
        t.append((time.perf_counter() - t0) * 1e6)
    assert rc == 0
    out["c_call_us_median"] = float(np.median(t))
    # the bare H2D of the same bytes, for scale
    t = []
    for _ in range(100):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        devbuf.copy_(pinned, non_blocking=True)
        torch.cuda.synchronize()
        t.append((time.perf_counter() - t0) * 1e6)
    out["bare_h2d_us_median"] = float(np.median(t))
    print(json.dumps(out))


if __name__ == "__main__":
    main()

#!/usr/bin/env python
"""A/B of fused-scan kernel builds: device time of the scan kernel (and of the whole per-motif
step: scan + site pass) for one representative motif per group count G = 2..8 over a synthetic
sequence, for every library given (CGBK_LIB selects the build; one subprocess per library).

    python tools/scan_variants.py [--bp 250000000] [--libs libcgbk.so,libcgbk_t512.so]

Variant libraries come from cgb_b200.build.build_library(out="libcgbk_x.so", defines=[...]); the
epilogue keeps three switches that rebuild earlier forms for such comparisons: -DCGBK_MINMAX_DIFF
(|f - r| as max - min), -DCGBK_F2I_CORR (soft-max correction through F2I), -DCGBK_BIN_MULHI (histogram
bin as subtract + high multiply); thread counts per group count: -DCGBK_STATS_THREADS_G3=640 etc.
"""
import argparse
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(bp, hits):
    import numpy as np
    import torch

    import bench
    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, _cabi, prepare_thresholds
    from cgb_b200 import motif_batch as mb
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)
    site_lists = bench.motif_site_lists(60)
    by_g = {}
    for s in site_lists:
        by_g.setdefault((len(s[0]) + 3) // 4, s)
    gs = sorted(by_g)
    models = [PSSMModel([SiteCollection(by_g[g])], [1.0], device=dev) for g in gs]
    batch = mb.ModelBatch(models, dev)
    prepare_thresholds(models)
    thr = [M.threshold() for M in models] if hits else None
    g = PackedGenome([bp], dev)
    g.pack_into(0, bench.generate_bases(0, bp, dev))
    torch.cuda.synchronize()
    cap = int(sum(mb.expected_hits(M, bp) * 1.25 + 4096 for M in models) * 1.05)
    out = {}
    _cabi.profile_enable(2)
    acc = np.zeros(len(models))
    reps = 5
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    for it in range(2 + reps):
        if it == 2:
            acc[:] = 0
            ev[0].record()
        mb.scan_batch(models, g, thr, n_bins=256, to_host=False, recycle=True, hit_cap=cap, kernel_ms=acc)
    ev[1].record()
    torch.cuda.synchronize()
    k = acc / reps
    for G, M, ms in zip(gs, models, k):
        out["G%d_m%d" % (G, M.length)] = round(float(ms), 4)
    out["sum_kernel_ms"] = round(float(k.sum()), 4)
    out["step_ms"] = round(ev[0].elapsed_time(ev[1]) / reps, 4)
    print(json.dumps(out))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=250_000_000)
    ap.add_argument("--libs", default="libcgbk.so")
    ap.add_argument("--no-hits", action="store_true")
    ap.add_argument("--child", action="store_true")
    a = ap.parse_args()
    if a.child:
        child(a.bp, not a.no_hits)
        return
    for lib in a.libs.split(","):
        env = dict(os.environ)
        env["CGBK_LIB"] = os.path.join(ROOT, "cgb_b200", "_lib", lib)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "--child", "--bp", str(a.bp)] +
                           (["--no-hits"] if a.no_hits else []), env=env, capture_output=True, text=True)
        line = [l for l in r.stdout.splitlines() if l.startswith("{")]
        print(lib, line[-1] if line else "FAILED: " + r.stderr[-400:])


if __name__ == "__main__":
    main()

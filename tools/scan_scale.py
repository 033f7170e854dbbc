#!/usr/bin/env python
"""Steady-state throughput of the tiled scan kernels on a large synthetic sequence.

    python tools/scan_scale.py [--bp 1000000000] [--reps 5] [--m 22]

Generates i.i.d. uniform DNA on the GPU (BASELINE configs[3] style), packs it, and
times (CUDA events around the tiled kernel, via cgbk_profile_*) the thresholded
hit scan (16-bit filter kernel) and the background-statistics scan (int32 kernel).
Prints one JSON object per mode.
"""
from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, _cabi

    ap = argparse.ArgumentParser()
    ap.add_argument("--bp", type=int, default=1_000_000_000)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--m", type=int, default=22)
    ap.add_argument("--plane", action="store_true", help="also write the score plane in the stats scan")
    ap.add_argument("--bins", type=int, default=256, help="histogram bins of the stats scan")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(dev)

    if args.m == 22:
        with open(os.path.join(ROOT, "tests", "golden", "lexa_sites.json")) as f:
            d = json.load(f)
        site_lists = [mo["sites"] for mo in d["motifs"]]
        counts = [len(s) for s in site_lists]
        M = PSSMModel([SiteCollection(s) for s in site_lists], [c / float(sum(counts)) for c in counts], device=dev)
    else:
        rng = np.random.default_rng(args.m)
        pwm = rng.dirichlet([0.5] * 4, size=args.m) * 0.92 + 0.02
        M = PSSMModel.from_pwm({l: pwm[:, k].tolist() for k, l in enumerate("ACGT")}, device=dev)

    gen = torch.Generator(device=dev)
    gen.manual_seed(4)
    g = PackedGenome([args.bp], dev)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device=dev)
    chunk = 1 << 28
    # pack in chunks of 2^28 bases (multiple of the 512-base alignment)
    ascii_all = torch.empty(args.bp, dtype=torch.uint8, device=dev)
    for off in range(0, args.bp, chunk):
        n = min(chunk, args.bp - off)
        ascii_all[off:off + n] = lut[torch.randint(0, 4, (n,), device=dev, generator=gen)]
    g.pack_into(0, ascii_all)
    torch.cuda.synchronize()
    del ascii_all

    nwin = g.total_windows(M.length)
    sms = torch.cuda.get_device_properties(dev).multi_processor_count
    issue_peak = sms * 128 * 1965e6
    _cabi.profile_enable(True)
    thr = M.threshold() if args.m == 22 else M.pssm.max - 0.15 * (M.pssm.max - M.pssm.min)

    out = {}
    for mode in ("hits", "stats"):
        times = []
        extra = {}
        for rep in range(args.reps + 2):
            if mode == "hits":
                res = M.scan_hits(g, thr)
                extra = dict(M.last_scan_counters)
            else:
                bg = M.background_stats(g, n_bins=args.bins, keep_scores=args.plane)
                torch.cuda.synchronize()
                st = bg["stats"].cpu().numpy()
                extra = {"n": float(st[0]), "mean": float(st[3]), "std": float(st[4])}
            if rep >= 2:
                times.append(_cabi.profile_last_ms())
        ms = float(np.mean(times))
        info = _cabi.last_launch_info()
        rate = nwin / (ms * 1e-3)
        out[mode] = {"mode": mode, "bp": args.bp, "m": M.length, "kernel_ms": ms, "bp_motif_per_s": rate,
                     "issue_frac": rate * 2 * M.length / issue_peak, "hbm_GBps": nwin * 0.25 / (ms * 1e-3) / 1e9,
                     "tiles": info["tiles"], "lane_bases": info["lane_bases"], **extra}
        print(json.dumps(out[mode]))


if __name__ == "__main__":
    main()

"""Diagnostic: are the background moments repeatable call to call, and equal between the histogram
scan (background_stats) and the moments-only batched scan?"""
import json, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
from cgb_b200 import PackedGenome, PSSMModel, SiteCollection
from cgb_b200 import motif_batch as mb
from oracle import np_oracle as O

d = json.load(open(os.path.join(ROOT, "tests", "golden", "lexa_sites.json")))
sites = [mo["sites"] for mo in d["motifs"]]
cnt = [len(s) for s in sites]
M = PSSMModel([SiteCollection(s) for s in sites], [c / float(sum(cnt)) for c in cnt])
for n, nf in ((2_000_000, 0.0005), (2_000_000, 0.0), (1_999_981, 0.0005), (40_000_000, 0.0005)):
    seq = O.synthetic_genome(n, 63, n_frac=nf)
    g = PackedGenome.from_sequences([seq])
    outs = []
    for rep in range(3):
        a = M.background_stats(g, n_bins=256)["stats"].cpu().numpy()[:3]
        b = mb.scan_batch([M], g, None, n_bins=0)["stats"][0][:3]
        c = mb.scan_batch([M], g, None, n_bins=256)["stats"][0][:3]
        outs.append((a, b, c))
    print(n, nf)
    for a, b, c in outs:
        print("  hist", [x.hex() for x in a.tolist()], "\n  mom ", [x.hex() for x in b.tolist()], "\n  bat ", [x.hex() for x in c.tolist()],
              "d(sum)=%g d(sumsq)=%g" % (a[1] - b[1], a[2] - b[2]))

#!/usr/bin/env python
"""Randomised parity run of the CUDA path against the C/numpy oracle (not a pytest: runs for a time
budget and reports the first mismatch with its seed).  Random motif lengths 1-32, several
sequences per set from a handful of bases to a few Mbp, N runs, lower-case stretches, random
thresholds, random promoter slices.

    python tools/fuzz_parity.py [--seconds 120] [--seed 0]
"""
import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def random_sequence(rng, n):
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    if n > 50 and rng.random() < 0.6:
        for _ in range(int(rng.integers(1, 6))):
            at = int(rng.integers(0, n - 10)); ln = int(rng.integers(1, 40))
            seq[at:at + ln] = ord("N") if rng.random() < 0.7 else ord(rng.choice(list("RYKMSWn")))
    if n > 200 and rng.random() < 0.4:
        at = int(rng.integers(0, n - 100)); ln = int(rng.integers(1, 100))
        seq[at:at + ln] |= 0x20                     # lower case
    return seq.tobytes()


def one_case(seed):
    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, _cabi
    from oracle import c_oracle, np_oracle as O
    rng = np.random.default_rng(seed)
    m = int(rng.integers(1, 33))
    cols = rng.dirichlet([rng.choice([0.3, 0.5, 1.0, 3.0])] * 4, size=m)
    sites = ["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m)) for _ in range(int(rng.integers(5, 40)))]
    M = PSSMModel([SiteCollection(sites)], [1.0])
    om = O.OracleModel([sites], [1.0])
    sizes = [int(rng.choice([m - 1 if m > 1 else 1, m, m + 1, 100, 1000, 1023, 1024, 1025, 5000, 33000]))
             for _ in range(int(rng.integers(0, 4)))]
    sizes.append(int(rng.choice([20000, 150000, 600000, 1500000, 2500000, 3600000])))
    rng.shuffle(sizes)
    seqs = [random_sequence(rng, n) for n in sizes]
    g = PackedGenome.from_sequences(seqs)
    info = {"seed": seed, "m": m, "sizes": sizes}
    # hits at a random operating point
    allb = np.concatenate([c_oracle.score_seq(om, s)[2] for s in seqs if len(s) >= m])
    thr = float(np.quantile(allb, rng.choice([0.9, 0.99, 0.999, 0.9999])))
    seq_i, pos, strand, score = M.scan_hits(g, thr)
    for k, s in enumerate(seqs):
        sel = seq_i == k
        if len(s) < m:
            assert not sel.any(), info
            continue
        ep, es, ev = c_oracle.scan_hits(om, s, thr)
        assert np.array_equal(pos[sel], ep) and np.array_equal(strand[sel], es) and np.array_equal(score[sel], ev), \
            (info, "hits", k)
    # background record
    bins = int(rng.choice([1, 64, 256, 300, 1024]))
    bg = M.background_stats(g, n_bins=bins, keep_scores=True)
    st = bg["stats"].cpu().numpy()
    assert st[_cabi.BG_N] == len(allb) and int(bg["hist"].sum()) == len(allb), (info, "count")
    assert abs(st[_cabi.BG_MEAN] - allb.mean()) < 1e-6 and abs(st[_cabi.BG_STD] - allb.std()) < 1e-6, (info, "moments")
    if allb.std() < 1e-3:
        return info
    # posteriors of random slices of the largest sequence, from the plane and re-scored
    M.fit_background(bg=bg)
    O.build_bayesian_estimator(om, allb)
    if om._std_m < 1e-3:
        # all sites identical: the reference's own posterior hinges on exact float equality of a
        # window score with the site mean (pdf of a zero-width normal); not a meaningful target
        return info
    k = int(np.argmax(sizes))
    n = sizes[k]
    starts = rng.integers(0, max(1, n - 2 * m - 2), size=40)
    ends = np.minimum(n, starts + rng.integers(m, min(n, 3000) + 1, size=40))
    keep = ends - starts >= m
    starts, ends = starts[keep], ends[keep]
    exp = np.array([c_oracle.binding_probability(om, seqs[k][int(a):int(b)], om._mu_m, om._std_m, om._mu_bg,
                                                 om._std_bg, 0.03, 1 / 350.0) for a, b in zip(starts, ends)])
    ok = np.isfinite(exp)
    if not ok.any():
        return info
    # the same with the GPU's own background estimate: separates the posterior kernel's error
    # from the sensitivity of a posterior to the (1e-6-accurate) background mean / std
    exp_own = np.array([c_oracle.binding_probability(om, seqs[k][int(a):int(b)], om._mu_m, om._std_m, M._mu_bg,
                                                     M._std_bg, 0.03, 1 / 350.0) for a, b in zip(starts, ends)])
    for use_plane in (True, False):
        got = M.binding_probabilities(g, starts, ends, 0.03, 1 / 350.0, seq_index=k, use_plane=use_plane).cpu().numpy()
        err = float(np.max(np.abs(got[ok] - exp[ok])))
        err_own = float(np.max(np.abs(got[ok] - exp_own[ok])))
        info["posterior_err_plane" if use_plane else "posterior_err_rescore"] = (err, err_own)
        assert err < 1e-6, (info, "posterior", use_plane, err, "kernel alone", err_own,
                            "mu_bg diff", M._mu_bg - om._mu_bg, "std_bg diff", M._std_bg - om._std_bg)
    return info


def sites_case(seed):
    """identify_sites / calculate_regulation_probabilities / regulation_pipeline on random gene
    tables (overlapping genes, regions running off either end, regions shorter than the motif)
    against the oracle's per-gene loops."""
    from cgb_b200 import (GeneTable, PackedGenome, PSSMModel, SiteCollection, calculate_regulation_probabilities,
                          identify_sites)
    from oracle import c_oracle, np_oracle as O
    rng = np.random.default_rng(seed)
    m = int(rng.integers(4, 31))
    cols = rng.dirichlet([0.5] * 4, size=m)
    sites = ["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m)) for _ in range(int(rng.integers(8, 40)))]
    M = PSSMModel([SiteCollection(sites)], [1.0])
    om = O.OracleModel([sites], [1.0])
    info = {"seed": seed, "m": m}
    if om.IC < 2.0:
        return info                                   # threshold near the bulk: millions of sites
    seqs, tables, ogenes = [], [], []
    for k in range(int(rng.integers(1, 4))):
        n = int(rng.integers(2000, 60000))
        seq = random_sequence(rng, n)
        genes, pos, strand = [], int(rng.integers(0, 200)), 1
        while pos + 60 < n:
            glen = int(rng.integers(30, 900))
            genes.append(O.OGene(pos, min(n, pos + glen), strand))
            pos += glen + int(rng.integers(-20, 400))
            pos = max(0, pos)
            if rng.random() < 0.4:
                strand = -strand
        seqs.append(seq); ogenes.append(genes)
        tables.append(GeneTable([x.start for x in genes], [x.end for x in genes], [x.strand for x in genes], n,
                                seq_index=k))
    g = PackedGenome.from_sequences(seqs)
    up = int(rng.choice([0, 50, 300, 1000])); dw = int(rng.integers(0, 80)); use_up = bool(rng.random() < 0.5)
    got = identify_sites(M, g, tables, up, dw, use_up, strict=False)
    want = []
    for k, (seq, genes) in enumerate(zip(seqs, ogenes)):
        want.extend((k, s) for s in O.identify_sites(om, seq, genes, up if use_up else None, dw, skip_short=True))
    want.sort(key=lambda t: t[1].score, reverse=True)
    assert len(got) == len(want), (info, len(got), len(want))
    for a, (k, b) in zip(got, want):
        assert (a.chromid, a.start, a.end, a.strand, float(a.score), a.gene) == \
            (k, b.start, b.end, b.strand, float(b.score), b.gene), (info, a, b)
    info["sites"] = len(got)
    # posteriors of every gene with a promoter at least one motif long + the one-call pipeline
    allb = np.concatenate([c_oracle.score_seq(om, s)[2] for s in seqs])
    M.fit_background(g)
    O.build_bayesian_estimator(om, allb)
    if om._std_m < 1e-3 or up + dw < m:
        return info
    for k, (seq, genes, gt) in enumerate(zip(seqs, ogenes, tables)):
        ps, pe = gt.promoter_region_location(up, dw)
        lens = np.array([len(seq[int(a):int(b)]) for a, b in zip(ps, pe)])
        if (lens < m).any():
            continue                                  # the reference raises for such a gene
        got_p = calculate_regulation_probabilities(M, g, gt, 0.03, 1 / 350.0, up, dw).cpu().numpy()
        exp = np.array([c_oracle.binding_probability(om, seq[int(a):int(b)], om._mu_m, om._std_m, om._mu_bg,
                                                     om._std_bg, 0.03, 1 / 350.0) for a, b in zip(ps, pe)])
        ok = np.isfinite(exp)
        assert np.max(np.abs(got_p[ok] - exp[ok]), initial=0.0) < 1e-6, (info, "posterior", k)
    # one-call pipeline on the first replicon alone (it refits the model's estimator to that
    # replicon's background, so it comes last)
    seq, gt = seqs[0], tables[0]
    ps, pe = gt.promoter_region_location(up, dw)
    if all(len(seq[int(a):int(b)]) >= m for a, b in zip(ps, pe)):
        b0 = c_oracle.score_seq(om, seq)[2]
        O.build_bayesian_estimator(om, b0)
        p2, st2, h2 = M.regulation_pipeline(seq, ps, pe, 0.03, 1 / 350.0)
        assert int(h2.sum()) == len(b0) and abs(st2[3] - b0.mean()) < 1e-6 and abs(st2[4] - b0.std()) < 1e-6, info
        exp = np.array([c_oracle.binding_probability(om, seq[int(a):int(b)], om._mu_m, om._std_m, om._mu_bg,
                                                     om._std_bg, 0.03, 1 / 350.0) for a, b in zip(ps, pe)])
        ok = np.isfinite(exp)
        assert np.max(np.abs(p2[ok] - exp[ok]), initial=0.0) < 1e-6, (info, "pipeline posterior")
    return info


def chunks_case(seed):
    """The chunked upload of cgbk_regulation_pipeline (tile-range scans as chunks arrive) against
    the resident whole-sequence path: random sizes, 3 or 4 chunks forced through the tuning
    knob, non-ACGT runs scattered around (some of them on the cuts)."""
    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, _cabi
    rng = np.random.default_rng(seed)
    m = int(rng.integers(2, 33))
    cols = rng.dirichlet([0.5] * 4, size=m)
    sites = ["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m)) for _ in range(20)]
    n = int(rng.choice([rng.integers(2100, 9000), rng.integers(9000, 200000), rng.integers(200000, 2600000)]))
    chunks = int(rng.choice([3, 4]))
    seq = rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), size=n)
    fr = (0.5, 0.8) if chunks == 3 else (0.35, 0.65, 0.85)
    for f in fr:
        cut = int(f * n) // 512 * 512
        if rng.random() < 0.7 and cut > 40:
            a = cut + int(rng.integers(-30, 30)); seq[max(0, a):a + int(rng.integers(1, 50))] = ord("N")
    for _ in range(int(rng.integers(0, 4))):
        a = int(rng.integers(0, n)); seq[a:a + int(rng.integers(1, 30))] = ord("n")
    seq = seq.tobytes()
    n_reg = int(rng.integers(0, 60))
    starts = np.sort(rng.integers(0, max(1, n - 2 * m - 400), size=n_reg)).astype(np.int64)
    ends = np.minimum(n, starts + rng.integers(m, 400 + m, size=n_reg)).astype(np.int64)
    os.environ["CGBK_PIPE_CHUNKS"] = str(chunks)
    try:
        M = PSSMModel([SiteCollection(sites)], [1.0])
        post, stats, hist = M.regulation_pipeline(seq, starts, ends, 0.03, 1 / 350.0)
    finally:
        os.environ.pop("CGBK_PIPE_CHUNKS", None)
    M2 = PSSMModel([SiteCollection(sites)], [1.0])
    g = PackedGenome.from_sequences([seq])
    bg = M2.background_stats(g, keep_scores=True)
    M2.fit_background(bg=bg)
    st2 = bg["stats"].cpu().numpy()
    info = {"seed": seed, "m": m, "n": n, "chunks": chunks, "regions": n_reg}
    assert np.array_equal(hist, bg["hist"].cpu().numpy().astype(np.uint64)), (info, "hist")
    assert stats[_cabi.BG_N] == st2[_cabi.BG_N] == n - m + 1, (info, "n")
    assert abs(stats[_cabi.BG_MEAN] - st2[_cabi.BG_MEAN]) < 1e-12 and abs(stats[_cabi.BG_STD] - st2[_cabi.BG_STD]) < 1e-12, info
    if n_reg and M2._std_m > 1e-3:
        post2 = M2.binding_probabilities(g, starts, ends, 0.03, 1 / 350.0).cpu().numpy()
        assert np.max(np.abs(post - post2)) < 1e-12, (info, "posterior", float(np.max(np.abs(post - post2))))
    return info


def batch_case(seed):
    """cgbk_scan_batch (fused scan in its lean modes + site pass) on a random motif set over a ragged
    sequence set: per motif the background record against the oracle's scores of every window, the
    ordered site list bit for bit against the oracle's hit loop, the split transfer form against the
    8-byte records, and the record of a second run against the first (repeatability)."""
    from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, _cabi
    from cgb_b200 import motif_batch as mb
    from oracle import c_oracle, np_oracle as O
    rng = np.random.default_rng(seed)
    K = int(rng.integers(1, 6))
    lens = [int(rng.integers(1, 33)) for _ in range(K)]
    site_sets = []
    for m in lens:
        cols = rng.dirichlet([rng.choice([0.3, 0.5, 1.0, 3.0])] * 4, size=m)
        site_sets.append(["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m))
                          for _ in range(int(rng.integers(5, 40)))])
    Ms = [PSSMModel([SiteCollection(s)], [1.0]) for s in site_sets]
    oms = [O.OracleModel([s], [1.0]) for s in site_sets]
    m_max = max(lens)
    sizes = [int(rng.choice([m_max, m_max + 1, m_max + 2, 100, 1000, 1023, 1024, 1025, 4063, 4064, 4065, 4066, 4067,
                             33000])) for _ in range(int(rng.integers(0, 4)))]
    sizes.append(int(rng.choice([20000, 150000, 600000, 1500000, 2600000])))
    sizes = [max(n, m_max) for n in sizes]
    rng.shuffle(sizes)
    seqs = [random_sequence(rng, n) for n in sizes]
    g = PackedGenome.from_sequences(seqs)
    info = {"seed": seed, "lens": lens, "sizes": sizes}
    scores = [np.concatenate([c_oracle.score_seq(om, s)[2] for s in seqs]) for om in oms]
    thr = [float(np.quantile(b, rng.choice([0.5, 0.9, 0.99, 0.999]))) if rng.random() < 0.7 else M.threshold()
           for b, M in zip(scores, Ms)]
    bins = int(rng.choice([0, 1, 64, 256]))
    res = mb.scan_batch(Ms, g, thr, n_bins=bins)
    again = mb.scan_batch(Ms, g, thr, n_bins=bins, split=True)
    assert np.array_equal(res["stats"][:, :5], again["stats"][:, :5]), (info, "repeat")
    off = np.asarray(g.seq_off, dtype=np.int64)
    for k, (M, om, b) in enumerate(zip(Ms, oms, scores)):
        st = res["stats"][k]
        assert st[_cabi.BG_N] == len(b), (info, "count", k)
        if bins:
            assert int(res["hist"][k].sum()) == len(b), (info, "hist sum", k)
        assert abs(st[_cabi.BG_MEAN] - b.mean()) < 1e-6 and abs(st[_cabi.BG_STD] - b.std()) < 1e-6, \
            (info, "moments", k, st[_cabi.BG_MEAN] - b.mean(), st[_cabi.BG_STD] - b.std())
        gp, strand, score = mb.unpack_hits(res["hits"][k])
        assert np.array_equal(again["hits"][k].records(), res["hits"][k]), (info, "split", k)
        ep, es, ev = [], [], []
        for j, s in enumerate(seqs):
            p, sd, v = c_oracle.scan_hits(om, s, thr[k])
            ep.append(p + off[j]); es.append(sd); ev.append(v)
        ep, es, ev = np.concatenate(ep), np.concatenate(es), np.concatenate(ev)
        assert np.array_equal(gp, ep) and np.array_equal(strand, es) and np.array_equal(score, ev), (info, "hits", k)
    return info


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=120.0)
    ap.add_argument("--seed", type=int, default=0)
    ap.add_argument("--mode", choices=["kernels", "sites", "chunks", "batch"], default="kernels")
    args = ap.parse_args()
    t0 = time.time()
    n, seed = 0, args.seed
    worst = {"plane": 0.0, "rescore": 0.0}
    while time.time() - t0 < args.seconds:
        info = {"kernels": one_case, "sites": sites_case, "chunks": chunks_case, "batch": batch_case}[args.mode](seed)
        worst["plane"] = max(worst["plane"], info.get("posterior_err_plane", (0.0, 0.0))[0])
        worst["rescore"] = max(worst["rescore"], info.get("posterior_err_rescore", (0.0, 0.0))[0])
        seed += 1
        n += 1
    print(json.dumps({"cases": n, "first_seed": args.seed, "last_seed": seed - 1, "mismatches": 0,
                      "worst_posterior_error": worst, "seconds": round(time.time() - t0, 1)}))


if __name__ == "__main__":
    main()

"""GPU parity tests (run on the B200 box): the CUDA path through the C ABI against
the committed golden vectors and the CPU oracle on the same seeded inputs.

Tolerances (BASELINE.json north_star): hit positions / strands / hit sets and the
float32 per-strand scores bit-exact; soft-max scores within 1e-4 absolute (observed
< 1e-9 against the oracle, < 5e-7 against the numpy-2 golden run); posteriors
within 1e-6; background mean / std within 1e-6 (they feed the posterior)."""
import numpy as np
import pytest
import torch

from cgb_b200 import (GeneTable, PackedGenome, PSSMModel, SiteCollection, _cabi,
                      calculate_regulation_probabilities, identify_sites)
from oracle import c_oracle, np_oracle as O

pytestmark = pytest.mark.gpu

SCORE_TOL = 1e-4
POST_TOL = 1e-6


def _collections(lexa):
    return [SiteCollection(mo["sites"], None, mo["name"]) for mo in lexa["motifs"]]


@pytest.fixture(scope="module")
def models(lexa, lexa_weights):
    out = {}
    for wname, w in lexa_weights.items():
        out[wname] = (PSSMModel(_collections(lexa), w),
                      O.OracleModel([mo["sites"] for mo in lexa["motifs"]], w))
    return out


def lexa_weights_site_count(lexa):
    counts = [len(mo["sites"]) for mo in lexa["motifs"]]
    return [c / float(sum(counts)) for c in counts]


def random_model(m, seed):
    rng = np.random.default_rng(seed)
    pwm = rng.dirichlet([0.5] * 4, size=m) * 0.92 + 0.02
    vals = {l: pwm[:, k].tolist() for k, l in enumerate("ACGT")}
    M = PSSMModel.from_pwm(vals)
    om = O.OracleModel.from_logodds(np.array(M.pssm.rows()))
    return M, om


# --------------------------------------------------------------------------- pack
def test_pack_planes():
    rng = np.random.default_rng(1)
    alphabet = np.frombuffer(b"ACGTacgtNnRYKM-*", dtype=np.uint8)
    for n in (1, 31, 32, 33, 511, 512, 513, 5000, 70001):
        probs = np.array([20] * 4 + [2] * 4 + [1] * 8, dtype=float)
        seq = alphabet[rng.choice(len(alphabet), size=n, p=probs / probs.sum())].tobytes()
        g = PackedGenome.from_sequences([seq, seq[: n // 2 + 1]])
        torch.cuda.synchronize()
        b2 = g.bases2.cpu().numpy().view(np.uint32)
        amb = g.amb.cpu().numpy().view(np.uint32)
        low = g.lower.cpu().numpy().view(np.uint32)
        for k, s in enumerate([seq, seq[: n // 2 + 1]]):
            arr = np.frombuffer(s, dtype=np.uint8)
            code = O._CODE[arr]
            p = np.arange(len(arr)) + g.seq_off[k]
            got = (b2[p >> 4] >> ((p & 15) * 2).astype(np.uint32)) & 3
            assert np.array_equal(got, np.where(code < 0, 0, code)), (n, k)
            got_amb = (amb[p >> 5] >> (p & 31).astype(np.uint32)) & 1
            assert np.array_equal(got_amb, (code < 0).astype(np.uint32)), (n, k)
            is_low = (code >= 0) & (arr >= 97)
            got_low = (low[p >> 5] >> (p & 31).astype(np.uint32)) & 1
            assert np.array_equal(got_low, is_low.astype(np.uint32)), (n, k)
        # padding stays zero
        end = g.seq_off[0] + n
        pad = np.arange(end, g.seq_off[1])
        assert not ((b2[pad >> 4] >> ((pad & 15) * 2).astype(np.uint32)) & 3).any()
        assert not ((amb[pad >> 5] >> (pad & 31).astype(np.uint32)) & 1).any()


# ---------------------------------------------------------------------- score_seq
@pytest.mark.parametrize("wname", ["equal", "site_count"])
def test_score_seq_golden(golden, models, wname):
    M, om = models[wname]
    rec = golden["models"][wname]
    for cname in ("random2k", "with_N", "lowercase"):
        case = rec["score_seq"][cname]
        fwd = M.score_seq(case["seq"], both=False)
        assert isinstance(fwd, np.ndarray) and fwd.dtype == np.float32
        assert [float(x) for x in fwd] == case["fwd_f32"], cname           # bit-exact float32
        both = M.score_seq(case["seq"])
        assert isinstance(both, list) and isinstance(both[0], float)
        assert len(both) == len(case["seq"]) - 22 + 1
        assert np.max(np.abs(np.array(both) - np.array(case["both"]))) < 5e-7   # numpy-2 golden
        oracle = O.score_seq(om, case["seq"], True)[0]                     # pinned-numpy semantics
        assert np.max(np.abs(np.array(both) - np.array(oracle))) < 1e-9
    for s, v in rec["score_seq"]["single"].items():
        got = M.score_seq(s)
        assert isinstance(got, list) and len(got) == 1
        assert abs(got[0] - v["both"]) < 5e-7
        f = M.score_seq(s, both=False)
        assert isinstance(f, list) and len(f) == 1 and float(f[0]) == v["fwd"]
    assert np.max(np.abs(np.array([x[0] for x in M.score_self()]) - np.array(rec["score_self"]))) < 5e-7


def test_score_seq_kat_and_errors(golden):
    M = PSSMModel([SiteCollection(["ACGT"])], [1.0])
    k = golden["kat_acgt"]
    assert float(M.score_seq("ACGT", both=False)[0]) == k["fwd"]
    assert abs(M.score_seq("ACGT")[0] - 3.7122876) < 1e-6
    assert M.score_seq("ACG") == []        # one base short: Biopython returns an empty array
    with pytest.raises(ValueError):
        M.score_seq("AC")                  # numpy.empty(len - m + 1) with a negative count raises
    assert M.score_seqs([]) == []


@pytest.mark.parametrize("m", [1, 3, 4, 5, 8, 13, 16, 17, 21, 24, 27, 30, 32])
def test_dense_scores_all_lengths(m):
    M, om = random_model(m, 100 + m)
    seq = O.synthetic_genome(3000, m, n_frac=0.01, n_run=3)
    f, r, b = c_oracle.score_seq(om, seq)
    assert np.array_equal(M.score_seq(seq, both=False), f)
    assert np.max(np.abs(np.array(M.score_seq(seq)) - b)) < 1e-9


def test_motif_too_long_is_rejected():
    M, _ = random_model(33, 1)
    with pytest.raises(NotImplementedError):
        M.score_seq("A" * 50)


# ---------------------------------------------------------------------- scan_hits
def _assert_hits_equal(M, om, seqs, thr):
    g = PackedGenome.from_sequences(seqs)
    seq_i, pos, strand, score = M.scan_hits(g, thr)
    exp = [c_oracle.scan_hits(om, s, thr) for s in seqs]
    e_seq = np.concatenate([np.full(len(e[0]), k) for k, e in enumerate(exp)])
    assert np.array_equal(seq_i, e_seq)
    assert np.array_equal(pos, np.concatenate([e[0] for e in exp]))
    assert np.array_equal(strand, np.concatenate([e[1] for e in exp]))
    assert np.array_equal(score, np.concatenate([e[2] for e in exp]))      # bit-exact float32
    return len(pos)


def test_scan_hits_lexa(models):
    M, om = models["site_count"]
    thr = M.threshold()
    seqs = [O.synthetic_genome(300000, 21), O.synthetic_genome(70001, 22, n_frac=0.003),
            O.synthetic_genome(40, 23), O.synthetic_genome(21, 24), O.synthetic_genome(22, 25)]
    # plant real sites so the hit list is not only background
    planted = bytearray(seqs[0])
    for k, site in enumerate(M.sites[:40]):
        planted[5000 * k + 17:5000 * k + 39] = site.encode()
        planted[5000 * k + 2500:5000 * k + 2522] = O.reverse_complement(site).encode()
    seqs[0] = bytes(planted)
    n = _assert_hits_equal(M, om, seqs, thr)
    assert n >= 80
    assert M.last_scan_counters["dirty_tiles"] > 0
    # lower threshold: many more hits; threshold above the maximum: none
    assert _assert_hits_equal(M, om, seqs, thr - 5.0) > 5 * n
    assert _assert_hits_equal(M, om, seqs, M.pssm.max + 1.0) == 0
    # minus-strand accumulation order of rev_comp_pssm (pssm_model.py:115)
    g = PackedGenome.from_sequences(seqs[:1])
    s2, p2, st2, sc2 = M.scan_hits(g, thr, revseq_order=False)
    f, r, _ = c_oracle.score_seq(om, seqs[0])
    assert np.array_equal(p2[st2 == 1], np.nonzero(f.astype(np.float64) >= thr)[0])
    assert np.array_equal(p2[st2 == -1], np.nonzero(r.astype(np.float64) >= thr)[0])
    assert np.array_equal(sc2[st2 == -1], r[p2[st2 == -1]])


def test_scan_hits_capacity_retry(models):
    M, om = models["equal"]
    seq = O.synthetic_genome(200000, 31)
    g = PackedGenome.from_sequences([seq])
    thr = M.threshold() - 6.0
    a = M.scan_hits(g, thr, hit_cap=1)
    e = c_oracle.scan_hits(om, seq, thr)
    assert len(e[0]) > 1000
    assert np.array_equal(a[1], e[0]) and np.array_equal(a[2], e[1]) and np.array_equal(a[3], e[2])


@pytest.mark.parametrize("m", [1, 2, 4, 6, 8, 11, 12, 15, 16, 19, 20, 22, 25, 28, 29, 32])
def test_scan_hits_all_lengths(m):
    M, om = random_model(m, 200 + m)
    seqs = [O.synthetic_genome(120000, 300 + m, n_frac=0.001, n_run=7), O.synthetic_genome(1500, m)]
    lo, hi = M.pssm.min, M.pssm.max
    for thr in (hi - 0.25 * (hi - lo), hi - 0.02 * (hi - lo)):
        _assert_hits_equal(M, om, seqs, thr)


def test_scan_hits_full_config1(models):
    """BASELINE configs[0]: LexA 22-bp PSWM over a synthetic 3.3 Mbp genome, both strands."""
    M, om = models["site_count"]
    seq = O.synthetic_genome(3309401, 1)
    n = _assert_hits_equal(M, om, [seq], M.threshold())
    assert 500 < n < 5000


def test_identify_sites_matches_reference_loop(models):
    M, om = models["site_count"]
    L = 120000
    seq = bytearray(O.synthetic_genome(L, 41, n_frac=0.002))
    for k, site in enumerate(M.sites[:30]):
        seq[4000 * k + 100:4000 * k + 122] = site.encode()
    seq = bytes(seq)
    genes = O.synthetic_genes(L, 60)
    # first gene reverse-strand near the start: `end - down` goes negative (Python slice quirk)
    genes[0] = O.OGene(2, 30, -1)
    gt = GeneTable([g.start for g in genes], [g.end for g in genes], [g.strand for g in genes], L)
    g = PackedGenome.from_sequences([seq])
    for up, use_up in ((300, False), (300, True)):
        # that first region is empty: the reference raises inside Biopython there, and so do the
        # product and the oracle unless told to leave such regions out
        with pytest.raises(ValueError):
            identify_sites(M, g, [gt], promoter_up_distance=up, promoter_dw_distance=50,
                           use_up_dist_site_scan=use_up, threshold=M.threshold() - 2.0)
        with pytest.raises(ValueError):
            O.identify_sites(om, seq.decode(), genes, up if use_up else None, 50, M.threshold() - 2.0)
        got = identify_sites(M, g, [gt], promoter_up_distance=up, promoter_dw_distance=50,
                             use_up_dist_site_scan=use_up, threshold=M.threshold() - 2.0, strict=False)
        exp = O.identify_sites(om, seq.decode(), genes, up if use_up else None, 50, M.threshold() - 2.0,
                               skip_short=True)
        assert len(exp) > 10
        assert [(s.start, s.end, s.strand, float(s.score), s.gene) for s in got] == \
               [(s.start, s.end, s.strand, float(s.score), s.gene) for s in exp]


# ------------------------------------------------------------------- background
def _oracle_bg(om, seqs):
    tot = np.zeros(3)
    scores = []
    for s in seqs:
        if len(s) >= om.length:
            _, _, b = c_oracle.score_seq(om, s)
            scores.append(b)
    b = np.concatenate(scores)
    return b


@pytest.mark.parametrize("m", [22, 8, 30])
def test_background_stats(models, m):
    if m == 22:
        M, om = models["site_count"]
    else:
        M, om = random_model(m, 500 + m)
    seqs = [O.synthetic_genome(400000, 51), O.synthetic_genome(90001, 52, n_frac=0.004), O.synthetic_genome(10, 53)]
    g = PackedGenome.from_sequences(seqs)
    bg = M.background_stats(g, n_bins=1024)
    stats = bg["stats"].cpu().numpy()
    hist = bg["hist"].cpu().numpy()
    b = _oracle_bg(om, seqs)
    assert stats[_cabi.BG_N] == len(b) == g.total_windows(m)
    assert abs(stats[_cabi.BG_MEAN] - np.mean(b)) < 1e-6
    assert abs(stats[_cabi.BG_STD] - np.std(b)) < 1e-6
    assert hist.sum() == len(b)
    edges = np.linspace(bg["hist_lo"], bg["hist_hi"], 1025)
    exp_hist, _ = np.histogram(b, bins=edges)
    # a score within float32 resolution of a bin edge may land in the neighbouring bin
    near_edge = np.abs(b[:, None] - edges[np.searchsorted(edges, b)[:, None] + np.array([-1, 0])]).min(axis=1) < 4e-6
    assert np.abs(np.cumsum(hist) - np.cumsum(exp_hist)).max() <= max(2, near_edge.sum())
    # accumulate a second genome into the same buffers
    g2 = PackedGenome.from_sequences([O.synthetic_genome(50000, 54)])
    M.background_stats(g2, n_bins=1024, accumulate_into=bg)
    b2 = np.concatenate([b, _oracle_bg(om, [O.synthetic_genome(50000, 54)])])
    stats = bg["stats"].cpu().numpy()
    assert stats[_cabi.BG_N] == len(b2)
    assert abs(stats[_cabi.BG_MEAN] - np.mean(b2)) < 1e-6 and abs(stats[_cabi.BG_STD] - np.std(b2)) < 1e-6


# -------------------------------------------------------------------- posteriors
@pytest.mark.parametrize("wname", ["equal", "site_count"])
def test_posteriors_golden(golden, models, wname):
    M, om = models[wname]
    rec = golden["models"][wname]
    seq = rec["score_seq"]["random2k"]["seq"]
    M.build_bayesian_estimator([[x] for x in M.score_seq(seq)])
    est = rec["estimator"]
    assert abs(M._mu_m - est["mu_m"]) < 1e-6 and abs(M._std_m - est["std_m"]) < 1e-6
    assert abs(M._mu_bg - est["mu_bg"]) < 1e-6 and abs(M._std_bg - est["std_bg"]) < 1e-6
    for p in rec["posteriors"]:
        got = M.binding_probability(p["seq"], p["p_motif"], p["alpha"])
        assert isinstance(got, float)
        assert abs(got - p["posterior"]) < POST_TOL
    p = rec["posteriors"][-1]
    assert abs(M.binding_probability(p["seq"], 0.01) - p["posterior"]) < POST_TOL   # default alpha 1/350
    with pytest.raises(ValueError):
        M.binding_probability("ACGT", 0.03)        # 4 < m - 1
    fresh = PSSMModel(M.site_collections, [1.0 / 6] * 6)
    with pytest.raises(AttributeError):
        fresh.binding_probability(seq[:350], 0.03)       # estimator not built, as in the reference


def test_config2_posteriors(models, lexa):
    """BASELINE configs[1] at reduced size for the test: background fit over all windows +
    posterior of every promoter, against the oracle."""
    M, om = models["site_count"]
    L, n_genes = 600000, 360
    seq = bytearray(O.synthetic_genome(L, 2, n_frac=0.0005))
    genes = O.synthetic_genes(L, n_genes)
    for k in range(0, n_genes, 9):            # plant sites in some promoters
        gene = genes[k]
        at = gene.start - 120 if gene.strand == 1 else gene.end + 60
        seq[at:at + 22] = M.sites[k % len(M.sites)].encode()
    seq = bytes(seq)
    g = PackedGenome.from_sequences([seq])
    gt = GeneTable([x.start for x in genes], [x.end for x in genes], [x.strand for x in genes], L)
    M.fit_background(g)
    b = _oracle_bg(om, [seq])
    O.build_bayesian_estimator(om, b)
    assert abs(M._mu_bg - om._mu_bg) < 1e-6 and abs(M._std_bg - om._std_bg) < 1e-6
    assert abs(M._mu_m - om._mu_m) < 1e-9 and abs(M._std_m - om._std_m) < 1e-9
    prior, alpha = lexa["prior_regulation_probability"], lexa["alpha"]
    post = calculate_regulation_probabilities(M, g, gt, prior, alpha, 300, 50).cpu().numpy()
    exp = []
    for gene in genes:
        s, e = O.promoter_region_location(gene, L, 300, 50)
        exp.append(c_oracle.binding_probability(om, seq[s:e], om._mu_m, om._std_m, om._mu_bg, om._std_bg,
                                                prior, alpha))
    exp = np.array(exp)
    assert (exp > 0.5).sum() >= 20
    assert np.max(np.abs(post - exp)) < POST_TOL          # scores read from the scan's score plane
    assert M._last_bg["plane"] is not None
    s, e = gt.promoter_region_location(300, 50)
    rescored = M.binding_probabilities(g, s, e, prior, alpha, use_plane=False).cpu().numpy()
    assert np.max(np.abs(rescored - exp)) < POST_TOL      # every window re-scored in float64
    # a hit scan between the two must not disturb the plane's tiling
    M.scan_hits(g)
    again = calculate_regulation_probabilities(M, g, gt, prior, alpha, 300, 50).cpu().numpy()
    assert np.array_equal(again, post)
    # the one-call host-buffer entry point (cgbk_regulation_pipeline): same numbers
    M2 = PSSMModel(M.site_collections, lexa_weights_site_count(lexa))
    p2, stats2, hist2 = M2.regulation_pipeline(seq, s, e, prior, alpha, n_bins=256)
    assert np.max(np.abs(p2 - exp)) < POST_TOL
    assert stats2[_cabi.BG_N] == len(b) and hist2.sum() == len(b)
    assert abs(M2._mu_bg - om._mu_bg) < 1e-6 and abs(M2._std_bg - om._std_bg) < 1e-6
    assert abs(M2.binding_probability(seq[1000:1350].decode(), prior, alpha) -
               M.binding_probability(seq[1000:1350].decode(), prior, alpha)) < 1e-9
    with pytest.raises(ValueError):
        M2.regulation_pipeline(seq, [0], [10], prior, alpha)       # slice shorter than the motif

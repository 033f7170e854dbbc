"""Batched models and the batched (fused) scan against the oracle (GPU parity tests).

cgbk_model_batch_create / cgbk_scan_batch / cgbk_sort_hits / cgbk_score_sites_host /
cgbk_background_stats_regions through their Python wrappers: per motif the background record
(moments, histogram) and the thresholded site list must equal what the reference's arithmetic
gives (C oracle / numpy oracle), for motif sets of mixed lengths (BASELINE configs[4]), with
non-ACGT runs, with a window cap (a rank's shard), with tiny buffers (the overflow path) and on a
second device.
"""
import numpy as np
import pytest
import torch

from cgb_b200 import GeneTable, PackedGenome, PSSMModel, SiteCollection, _cabi
from cgb_b200 import motif_batch as mb
from oracle import c_oracle, np_oracle as O

pytestmark = pytest.mark.gpu


def random_motifs(n, seed, lo=8, hi=30, n_sites=20):
    """JASPAR-scale synthetic motifs as SURVEY.md 8d describes config 5: length ~U{lo..hi}, sites
    sampled from a Dirichlet(0.5)-column PWM, pseudocount 1."""
    rng = np.random.default_rng(seed)
    out = []
    for _ in range(n):
        m = int(rng.integers(lo, hi + 1))
        cols = rng.dirichlet([0.5] * 4, size=m)
        out.append(["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m)) for _ in range(n_sites)])
    return out


def models_of(site_lists):
    return ([PSSMModel([SiteCollection(s)], [1.0]) for s in site_lists],
            [O.OracleModel([s], [1.0]) for s in site_lists])


def check_against_oracle(res, k, om, seq, thr, n_bins, nwin_cap=None, tol=1e-6):
    """model k of a scan_batch result vs the oracle on `seq` (windows capped at nwin_cap)."""
    m = om.length
    sub = seq if nwin_cap is None else seq[:nwin_cap + m - 1]
    _, _, b = c_oracle.score_seq(om, sub)
    st = res["stats"][k]
    assert st[_cabi.BG_N] == len(b)
    assert abs(st[_cabi.BG_MEAN] - b.mean()) < tol and abs(st[_cabi.BG_STD] - b.std()) < tol
    if res["hist"] is not None:
        h = res["hist"][k]
        assert h.sum() == len(b)
        lo, hi = np.floor(O.pssm_min(om.pssm)), np.ceil(O.pssm_max(om.pssm)) + 1.0
        eh, _ = np.histogram(b, bins=np.linspace(lo, hi, n_bins + 1))
        # fixed-point scores sit within ~1e-6 of the float64 ones: a score that close to a bin edge
        # may fall on the other side
        assert np.abs(np.cumsum(h) - np.cumsum(eh)).max() <= max(3, len(b) // 100000)
    if thr is not None:
        gpos, strand, score = mb.unpack_hits(res["hits"][k])
        ep, es, ev = c_oracle.scan_hits(om, sub, thr)
        assert np.array_equal(gpos, ep) and np.array_equal(strand, es) and np.array_equal(score, ev)
        assert res["counters"][k][_cabi.CTR_HITS] == len(ep)


@pytest.fixture(scope="module")
def motif_set():
    site_lists = random_motifs(26, 5)
    # the extremes of the supported range ride along: 1-base and 32-base motifs
    rng = np.random.default_rng(1)
    site_lists.append(["ACGT"[i] for i in rng.integers(0, 4, 12)])
    site_lists.append(["".join("ACGT"[i] for i in rng.integers(0, 4, 32)) for _ in range(9)])
    return site_lists


def test_model_batch_and_host_self_scores(motif_set):
    Ms, oms = models_of(motif_set)
    batch = mb.ModelBatch(Ms)
    assert all(M._handle is not None and M._handle.value for M in Ms)
    assert len({M._handle.value for M in Ms}) == len(Ms)
    for M, om in zip(Ms, oms):
        host = M.score_self()                                   # cgbk_score_sites_host
        dev = M.score_seqs(M.sites, True)                       # exact kernel
        exp = O.score_self(om)
        assert len(host) == len(exp)
        for a, b, c in zip(host, dev, exp):
            assert len(a) == 1 and abs(a[0] - c[0]) < 1e-12 and abs(b[0] - c[0]) < 1e-9
    del batch


def test_host_self_scores_ambiguous_and_lower_case(lexa, lexa_weights):
    site_lists = [list(mo["sites"]) for mo in lexa["motifs"]]
    site_lists[0][0] = site_lists[0][0][:5] + "N" + site_lists[0][0][6:]
    site_lists[0][1] = site_lists[0][1].lower()
    site_lists[1][0] = site_lists[1][0][:3] + "nR" + site_lists[1][0][5:10].lower() + site_lists[1][0][10:]
    M = PSSMModel([SiteCollection(s) for s in lexa_sites_upper(site_lists)], lexa_weights["site_count"])
    om = O.OracleModel(lexa_sites_upper(site_lists), lexa_weights["site_count"])
    # same model, but the sites scored are the edited ones
    M.__dict__["sites"] = [s for sl in site_lists for s in sl]
    exp = [O.score_seq(om, s, True)[0] for s in M.sites]
    got = M.score_self()
    for a, c in zip(got, exp):
        assert abs(a[0] - c[0]) < 1e-12


def lexa_sites_upper(site_lists):
    # PWM counts need clean ACGT sites
    return [[s.upper().replace("N", "A").replace("R", "A") for s in sl] for sl in site_lists]


@pytest.mark.parametrize("n_bins", [256, 0])
def test_scan_batch_mixed_lengths_vs_oracle(motif_set, n_bins):
    Ms, oms = models_of(motif_set)
    seq = O.synthetic_genome(400_003, 55, n_frac=0.001)
    g = PackedGenome.from_sequences([seq])
    thr = [M.threshold() for M in Ms]
    res = mb.scan_batch(Ms, g, thr, n_bins=n_bins)
    assert res["launches"] <= 4 * len(Ms)
    for k, om in enumerate(oms):
        check_against_oracle(res, k, om, seq, thr[k], n_bins)
    # background records only (no thresholds): same moments, no hit list
    res2 = mb.scan_batch(Ms, g, None, n_bins=n_bins)
    assert res2["hits"] is None
    assert np.array_equal(res2["stats"][:, :3], res["stats"][:, :3])


def test_scan_batch_window_cap_is_a_shard(motif_set):
    """A rank's shard: the buffer carries the halo of the longest motif, every motif stops at the
    shard's own windows -- the union over shards is the whole sequence, each window once."""
    Ms, oms = models_of(motif_set[:12])
    n, world = 300_011, 3
    seq = O.synthetic_genome(n, 56, n_frac=0.0005)
    m_max = max(M.length for M in Ms)
    thr = [M.threshold() - 1.0 for M in Ms]
    own = (n + world - 1) // world
    total = None
    hits = [[] for _ in Ms]
    for r in range(world):
        lo = r * own
        cap = min(own, n - lo)
        g = PackedGenome.from_sequences([seq[lo:min(n, lo + cap + m_max - 1)]])
        res = mb.scan_batch(Ms, g, thr, n_bins=128, win_cap=cap)
        for k, om in enumerate(oms):
            # the capped count also ends where the sequence ends
            assert res["stats"][k][_cabi.BG_N] == max(0, min(cap, n - lo - om.length + 1))
            gp, st, sc = mb.unpack_hits(res["hits"][k])
            hits[k].append((gp + lo, st, sc))
        part = np.concatenate([res["stats"][:, :3], res["hist"].astype(np.float64)], axis=1)
        total = part if total is None else total + part      # plays the all-reduce (sum)
    for k, om in enumerate(oms):
        _, _, b = c_oracle.score_seq(om, seq)
        nn, s1, s2 = total[k, 0], total[k, 1], total[k, 2]
        assert nn == len(b) and total[k, 3:].sum() == len(b)
        mean = s1 / nn
        assert abs(mean - b.mean()) < 1e-6 and abs(np.sqrt(max(s2 / nn - mean * mean, 0)) - b.std()) < 1e-6
        gp = np.concatenate([h[0] for h in hits[k]]); st = np.concatenate([h[1] for h in hits[k]])
        sc = np.concatenate([h[2] for h in hits[k]])
        ep, es, ev = c_oracle.scan_hits(om, seq, thr[k])
        assert np.array_equal(gp, ep) and np.array_equal(st, es) and np.array_equal(sc, ev)


def test_scan_batch_overflow_path_and_device_resident(motif_set):
    Ms, oms = models_of(motif_set[:8])
    seq = O.synthetic_genome(250_000, 57)
    g = PackedGenome.from_sequences([seq])
    thr = [M.threshold() - 2.0 for M in Ms]
    ref = mb.scan_batch(Ms, g, thr, n_bins=64)
    # a hit buffer far too small: sub-batches overflow, are cut, re-run and grown
    g2 = PackedGenome.from_sequences([seq])
    small = mb.scan_batch(Ms, g2, thr, n_bins=64, hit_cap=1024)
    for k in range(len(Ms)):
        assert np.array_equal(ref["hits"][k], small["hits"][k])
    assert np.array_equal(ref["stats"], small["stats"]) and np.array_equal(ref["hist"], small["hist"])
    # device-resident result (the benchmark's `value` leg): sorted on the device
    g3 = PackedGenome.from_sequences([seq])
    dev = mb.scan_batch(Ms, g3, thr, n_bins=64, to_host=False)
    for k in range(len(Ms)):
        assert np.array_equal(dev["hits"][k].cpu().numpy().view(np.uint64), ref["hits"][k])
    assert np.array_equal(dev["stats"].cpu().numpy()[:, :3], ref["stats"][:, :3])
    with pytest.raises(ValueError):
        mb.scan_batch(Ms, PackedGenome.from_sequences([seq]), thr, n_bins=64, to_host=False, hit_cap=1024)
    # the sink receives views of the pinned staging buffer
    got = {}
    mb.scan_batch(Ms, g, thr, n_bins=64, sink=lambda k, rec: got.__setitem__(k, rec.copy()))
    for k in range(len(Ms)):
        assert np.array_equal(got[k], ref["hits"][k])


def test_scan_batch_many_sequences_and_large_segments(lexa, lexa_weights):
    """A ragged multi-replicon buffer (device sequence tables) and a hit list long enough for the
    device radix sort."""
    site_lists = [mo["sites"] for mo in lexa["motifs"]]
    M = PSSMModel([SiteCollection(s) for s in site_lists], lexa_weights["site_count"])
    om = O.OracleModel(site_lists, lexa_weights["site_count"])
    rng = np.random.default_rng(8)
    seqs = [O.synthetic_genome(int(n), 60 + i, n_frac=0.002 if i % 2 else 0.0)
            for i, n in enumerate([50_000, 21, 22, 23, 0, 140_001, 3_000, 700_000])]
    g = PackedGenome.from_sequences(seqs)
    thr = M.threshold() - 10.5                      # a few percent of the windows: > 2^15 sites
    res = mb.scan_batch([M], g, [thr], n_bins=256)
    seq_i, pos, strand, score = mb.hits_of(g, res["hits"][0])
    assert len(pos) > (1 << 15)
    ep, es, ev, ei = [], [], [], []
    nwin = 0
    bs = []
    for i, s in enumerate(seqs):
        if len(s) < om.length:
            continue
        p, st, v = c_oracle.scan_hits(om, s, thr)
        ep.append(p); es.append(st); ev.append(v); ei.append(np.full(len(p), i))
        bs.append(c_oracle.score_seq(om, s)[2])
    b = np.concatenate(bs)
    assert np.array_equal(seq_i, np.concatenate(ei)) and np.array_equal(pos, np.concatenate(ep))
    assert np.array_equal(strand, np.concatenate(es)) and np.array_equal(score, np.concatenate(ev))
    st = res["stats"][0]
    assert st[_cabi.BG_N] == len(b) and abs(st[_cabi.BG_MEAN] - b.mean()) < 1e-6 and abs(st[_cabi.BG_STD] - b.std()) < 1e-6
    # the sort entry point on its own: a shuffled copy comes back ordered
    rec = torch.from_numpy(res["hits"][0].view(np.int64).copy()).cuda()
    perm = torch.randperm(rec.numel(), device="cuda")
    shuffled = rec[perm].contiguous()
    tmp = torch.empty(int(_cabi.lib().cgbk_sort_hits_workspace_bytes(rec.numel())), dtype=torch.uint8, device="cuda")
    _cabi.check(_cabi.lib().cgbk_sort_hits(shuffled.data_ptr(), rec.numel(), tmp.data_ptr(), tmp.numel(),
                                           torch.cuda.current_stream().cuda_stream))
    assert torch.equal(shuffled, rec)


def test_fused_scan_equals_filter_scan(lexa, lexa_weights):
    """The two hit paths (16-bit filter kernel of cgbk_scan_hits, fused fixed-point scan of
    cgbk_scan_batch) feed the same exact re-score: identical site lists, both strand orders."""
    site_lists = [mo["sites"] for mo in lexa["motifs"]]
    M = PSSMModel([SiteCollection(s) for s in site_lists], lexa_weights["equal"])
    seq = O.synthetic_genome(1_200_000, 61, n_frac=0.001)
    g = PackedGenome.from_sequences([seq])
    for revseq in (True, False):
        for dthr in (0.0, -4.0):
            thr = M.threshold() + dthr
            _, pos, strand, score = M.scan_hits(g, thr, revseq_order=revseq)
            res = mb.scan_batch([M], g, [thr], n_bins=0, revseq_order=revseq)
            gp, st, sc = mb.unpack_hits(res["hits"][0])
            assert np.array_equal(gp, pos) and np.array_equal(st, strand) and np.array_equal(sc, score)


def test_background_over_region_table_vs_oracle(lexa, lexa_weights):
    """cgbk_background_stats_regions: the reference's population (genome.py:215-218) -- upstream
    regions that overlap, repeat, run off both ends and hold N runs -- against the oracle."""
    site_lists = [mo["sites"] for mo in lexa["motifs"]]
    M = PSSMModel([SiteCollection(s) for s in site_lists], lexa_weights["site_count"])
    om = O.OracleModel(site_lists, lexa_weights["site_count"])
    n = 120_000
    seq = O.synthetic_genome(n, 62, n_frac=0.003)
    genes = O.synthetic_genes(n, 70)
    gt = GeneTable([x.start for x in genes], [x.end for x in genes], [x.strand for x in genes], n)
    g = PackedGenome.from_sequences([b"ACGT" * 100, seq])          # replicon 1 of a two-replicon buffer
    start, end = gt.upstream_noncoding_region_location(None, 50)
    bg = M.background_stats(g, regions=[(1, start, end)], n_bins=128)
    exp = np.array(O.upstream_background_scores(om, seq, genes, 50))
    st = bg["stats"].cpu().numpy()
    assert st[_cabi.BG_N] == len(exp) and int(bg["hist"].sum().item()) == len(exp)
    assert abs(st[_cabi.BG_MEAN] - exp.mean()) < 1e-9 and abs(st[_cabi.BG_STD] - exp.std()) < 1e-9
    lo, hi = M._score_bounds
    eh, _ = np.histogram(exp, bins=np.linspace(lo, hi, 129))
    assert np.abs(np.cumsum(bg["hist"].cpu().numpy()) - np.cumsum(eh)).max() <= 1
    # every window instead of all but the last; accumulation; an empty table
    bg2 = M.background_stats(g, regions=[(1, start, end)], n_bins=128, drop_last=False)
    from cgb_b200 import python_slice_bounds
    s_, e_ = python_slice_bounds(start, end, n)
    assert bg2["stats"][_cabi.BG_N].item() == int(np.maximum(e_ - s_ - om.length + 1, 0).sum())
    assert len(exp) == int(np.maximum(e_ - s_ - om.length, 0).sum())
    bg3 = M.background_stats(g, regions=[(1, start[:30], end[:30])], n_bins=128)
    bg3 = M.background_stats(g, regions=[(1, start[30:], end[30:])], n_bins=128, accumulate_into=bg3)
    assert torch.equal(bg3["hist"], bg["hist"]) and abs(bg3["stats"][_cabi.BG_MEAN].item() - st[_cabi.BG_MEAN]) < 1e-12
    empty = M.background_stats(g, regions=[(1, start[:0], end[:0])])
    assert empty["stats"][_cabi.BG_N].item() == 0 and np.isnan(empty["stats"][_cabi.BG_MEAN].item())
    # fit_background(regions=...) is build_bayesian_estimator over that population
    M.fit_background(g, regions=[(1, start, end)])
    assert abs(M._mu_bg - exp.mean()) < 1e-9 and abs(M._std_bg - exp.std()) < 1e-9


def test_moments_only_background_equals_histogram_scan(lexa, lexa_weights):
    site_lists = [mo["sites"] for mo in lexa["motifs"]]
    M = PSSMModel([SiteCollection(s) for s in site_lists], lexa_weights["site_count"])
    seq = O.synthetic_genome(2_000_000, 63, n_frac=0.0005)
    g = PackedGenome.from_sequences([seq])
    a = M.background_stats(g, n_bins=256)["stats"].cpu().numpy()
    res = mb.scan_batch([M], g, None, n_bins=0)                # MODE 2: no histogram
    assert np.array_equal(res["stats"][0][:5], a[:5])          # integer moments: identical


def test_background_record_is_repeatable_under_dynamic_tile_assignment(lexa, lexa_weights):
    """12 Mbp with non-ACGT runs: more tiles than resident warps, so the tiles fall to the warps in a
    different order every launch.  The moments are exact integer sums (table path and exactly scored
    tiles alike), so n / sum / sumsq / mean / std come out bit-identical launch after launch and
    between the histogram scan and the moments-only scan."""
    site_lists = [mo["sites"] for mo in lexa["motifs"]]
    M = PSSMModel([SiteCollection(s) for s in site_lists], lexa_weights["site_count"])
    seq = O.synthetic_genome(12_000_003, 66, n_frac=0.0005)
    g = PackedGenome.from_sequences([seq])
    first = M.background_stats(g, n_bins=256)
    a = first["stats"].cpu().numpy()[:5]
    for _ in range(3):
        again = M.background_stats(g, n_bins=256)
        assert np.array_equal(again["stats"].cpu().numpy()[:5], a) and torch.equal(again["hist"], first["hist"])
        assert np.array_equal(mb.scan_batch([M], g, None, n_bins=0)["stats"][0][:5], a)
        assert np.array_equal(mb.scan_batch([M], g, "patser", n_bins=256)["stats"][0][:5], a)


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_models_on_two_devices_from_one_process(lexa, lexa_weights):
    """The shared-memory attribute of the scan kernels and the SM count are per device: models
    on two GPUs of one process, neither of them on the current device for part of the calls."""
    site_lists = [mo["sites"] for mo in lexa["motifs"]]
    om = O.OracleModel(site_lists, lexa_weights["site_count"])
    seq = O.synthetic_genome(300_000, 64, n_frac=0.001)
    _, _, b = c_oracle.score_seq(om, seq)
    ep, es, ev = c_oracle.scan_hits(om, seq, om.threshold())
    for dev_index in (1, 0, 1):
        dev = torch.device("cuda", dev_index)
        M = PSSMModel([SiteCollection(s) for s in site_lists], lexa_weights["site_count"], device=dev)
        with torch.cuda.device(dev):
            g = PackedGenome.from_sequences([seq], dev)
        torch.cuda.set_device(0)                      # the current device is NOT always the model's
        with torch.cuda.device(dev):                  # (torch needs it for its own stream look-up)
            M.fit_background(g)
            _, pos, strand, score = M.scan_hits(g)
            res = mb.scan_batch([M], g, [M.threshold()])
        assert abs(M._mu_bg - b.mean()) < 1e-6 and abs(M._std_bg - b.std()) < 1e-6
        assert np.array_equal(pos, ep) and np.array_equal(strand, es) and np.array_equal(score, ev)
        gp, st, sc = mb.unpack_hits(res["hits"][0])
        assert np.array_equal(gp, ep) and np.array_equal(sc, ev)


def test_histogram_modes_agree_and_narrow_range_clamps(lexa, lexa_weights):
    """The lean scan (no score plane, default range: guard rows instead of clamps) and the generic
    one (score plane kept) must fill identical histograms and moments; a caller-chosen range
    narrower than the scores clamps: everything below the lower edge lands in the first bin,
    everything at or above the upper edge in the last (np.histogram of the clipped scores)."""
    site_lists = [mo["sites"] for mo in lexa["motifs"]]
    M = PSSMModel([SiteCollection(s) for s in site_lists], lexa_weights["site_count"])
    om = O.OracleModel(site_lists, lexa_weights["site_count"])
    seq = O.synthetic_genome(700_000, 65, n_frac=0.0008)
    g = PackedGenome.from_sequences([seq])
    lean = M.background_stats(g, n_bins=256)                        # no plane: lean kernel
    full = M.background_stats(g, n_bins=256, keep_scores=True)      # plane: generic kernel
    assert torch.equal(lean["hist"], full["hist"])
    assert np.array_equal(lean["stats"].cpu().numpy()[:5], full["stats"].cpu().numpy()[:5])
    one = mb.scan_batch([M], g, None, n_bins=256)                   # the batched entry point, lean as well
    assert np.array_equal(one["hist"][0], lean["hist"].cpu().numpy())
    _, _, b = c_oracle.score_seq(om, seq)
    mu = float(b.mean())
    lo, hi, nb = mu - 2.0, mu + 3.0, 37
    narrow = M.background_stats(g, n_bins=nb, hist_lo=lo, hist_hi=hi)
    h = narrow["hist"].cpu().numpy()
    assert h.sum() == len(b)
    edges = np.linspace(lo, hi, nb + 1)
    eh, _ = np.histogram(np.clip(b, lo, np.nextafter(hi, lo)), bins=edges)
    near = int((np.abs(b[:, None] - edges[None, 1:-1]).min(axis=1) < 1e-5).sum()) if len(b) < 2_000_000 else 8
    assert np.abs(np.cumsum(h) - np.cumsum(eh)).max() <= max(2, near)
    assert h[0] > 0.1 * len(b) and h[-1] > 0.001 * len(b)          # the clamped tails are really there
    assert abs(narrow["stats"][_cabi.BG_MEAN].item() - mu) < 1e-6  # moments do not depend on the range


def test_split_site_transfer_equals_the_8_byte_records(lexa, lexa_weights):
    """scan_batch(split=True): the sites cross PCIe as uint16 block-local offset + strand, float32
    score and one int32 per 4 096-base block; rebuilt on the host they are bit for bit the 8-byte
    records of the standard path -- two replicons (the second starts at a 512-aligned offset of the
    packed buffer), a dense, a sparse and a hit-less model, sub-batches forced by a small hit buffer,
    and the sink form."""
    rng = np.random.default_rng(21)
    site_lists = [["".join("ACGT"[i] for i in rng.integers(0, 4, m)) for _ in range(10)] for m in (8, 13, 22, 30)]
    Ms = [PSSMModel([SiteCollection(s)], [1.0]) for s in site_lists]
    Ms.append(PSSMModel([SiteCollection(s) for s in [mo["sites"] for mo in lexa["motifs"]]], lexa_weights["site_count"]))
    seqs = [O.synthetic_genome(700_001, 71, n_frac=0.0004), O.synthetic_genome(123_457, 72)]
    g = PackedGenome.from_sequences(seqs)
    thr = [M.threshold() for M in Ms]
    thr[2] = 1e9                                                   # no site at all
    ref = mb.scan_batch(Ms, g, thr, n_bins=64)
    got = mb.scan_batch(Ms, g, thr, n_bins=64, split=True)
    assert np.array_equal(ref["stats"], got["stats"]) and np.array_equal(ref["hist"], got["hist"])
    assert len(ref["hits"][2]) == 0 and len(got["hits"][2]) == 0
    for a, b in zip(ref["hits"], got["hits"]):
        assert isinstance(b, mb.SplitSites) and len(a) == len(b)
        assert np.array_equal(b.records(), a)
        gp, st, sc = mb.unpack_hits(a)
        assert np.array_equal(b.positions(), gp) and np.array_equal(b.strands(), st) and np.array_equal(b.score, sc)
        assert b.block_start[0] == 0 and b.block_start[-1] == len(a) and np.all(np.diff(b.block_start) >= 0)
    # small hit buffer -> several sub-batches; sink receives views of the staging memory
    seen = {}
    small = mb.scan_batch(Ms, g, thr, n_bins=64, split=True, hit_cap=max(len(h) for h in ref["hits"]) + 8,
                          sink=lambda k, s: seen.__setitem__(k, s.records().copy()))
    assert [small["hits"][k] for k in range(len(Ms))] == [len(h) for h in ref["hits"]]
    assert all(np.array_equal(seen[k], ref["hits"][k]) for k in range(len(Ms)))


def test_interleaved_processing_order_returns_the_callers_indices():
    """scan_batch processes a large batch in interleave_order (dense and sparse motifs alternate so
    that the site copies overlap the scans); records, histograms, site lists and the sink's indices
    must be those of the caller's order, identical to the in-order run."""
    rng = np.random.default_rng(33)
    lens = [8, 30, 12, 9, 22, 16, 8, 27, 10, 19]
    site_lists = [["".join("ACGT"[i] for i in rng.integers(0, 4, m)) for _ in range(int(rng.integers(4, 30)))] for m in lens]
    Ms = [PSSMModel([SiteCollection(s)], [1.0]) for s in site_lists]
    g = PackedGenome.from_sequences([O.synthetic_genome(900_000, 81, n_frac=0.0003)])
    a = mb.scan_batch(Ms, g, "patser", n_bins=32, interleave=False)
    seen = {}
    cap = int(max(len(h) for h in a["hits"]) * 1.2) + 64
    b = mb.scan_batch(Ms, g, "patser", n_bins=32, interleave=True, hit_cap=cap,
                      sink=lambda k, rec: seen.__setitem__(k, rec.copy()))
    assert sorted(b["order"]) == list(range(len(Ms))) and b["order"] != list(range(len(Ms)))
    assert np.array_equal(a["stats"], b["stats"]) and np.array_equal(a["hist"], b["hist"])
    assert np.array_equal(a["counters"][:, _cabi.CTR_HITS], b["counters"][:, _cabi.CTR_HITS])
    for k in range(len(Ms)):
        assert b["hits"][k] == len(a["hits"][k]) and np.array_equal(seen[k], a["hits"][k])
    c = mb.scan_batch(Ms, g, "patser", n_bins=32)                # default for >= 8 models to the host: interleaved
    assert all(np.array_equal(x, y) for x, y in zip(a["hits"], c["hits"]))


def test_underestimated_site_counts_rerun_with_larger_buffers(lexa, lexa_weights):
    """Explicit thresholds far below the Patser point: the expected site counts (guessed from the
    information content) are too small, the candidate buffer overflows in the MIDDLE of a sub-batch
    while the sites of the previous models are still on their way to the host, and the scan grows
    its buffers and runs the rest again.  Found by tools/fuzz_parity.py --mode batch (the in-flight
    copy used to be looked up in the re-allocated buffers)."""
    rng = np.random.default_rng(44)
    site_lists = [["".join("ACGT"[i] for i in rng.integers(0, 4, m)) for _ in range(12)] for m in (14, 10, 21, 9)]
    Ms = [PSSMModel([SiteCollection(s)], [1.0]) for s in site_lists]
    oms = [O.OracleModel([s], [1.0]) for s in site_lists]
    seq = O.synthetic_genome(400_000, 91, n_frac=0.0005)
    g = PackedGenome.from_sequences([seq])
    thr = [M.threshold() for M in Ms]
    thr[1] = float(Ms[1].pssm.min) - 1.0            # every window, both strands
    thr[3] = float(np.quantile(c_oracle.score_seq(oms[3], seq)[2], 0.5))
    mb.release_buffers()
    for split in (False, True):
        res = mb.scan_batch(Ms, g, thr, n_bins=32, split=split)
        for k, om in enumerate(oms):
            ep, es, ev = c_oracle.scan_hits(om, seq, thr[k])
            rec = res["hits"][k].records() if split else res["hits"][k]
            gp, st, sc = mb.unpack_hits(rec)
            assert np.array_equal(gp, ep) and np.array_equal(st, es) and np.array_equal(sc, ev), (split, k)
        assert len(res["hits"][1]) >= 2 * (len(seq) - 10 + 1) - 2000     # (windows with N are repaired, not skipped)
        mb.release_buffers()

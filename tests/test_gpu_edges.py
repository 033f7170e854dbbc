"""Edge cases of the CUDA path against the oracle: ragged sequence sets (shorter than, equal to and
just above the motif length), all-ambiguous and all-lower-case sequences, promoter slices with
negative / overshooting bounds, histogram bin counts at both ends, an empty promoter list."""
import numpy as np
import pytest

from cgb_b200 import PackedGenome, PSSMModel, SiteCollection, _cabi
from oracle import c_oracle, np_oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pair(lexa, lexa_weights):
    M = PSSMModel([SiteCollection(mo["sites"], None, mo["name"]) for mo in lexa["motifs"]], lexa_weights["site_count"])
    om = O.OracleModel([mo["sites"] for mo in lexa["motifs"]], lexa_weights["site_count"])
    return M, om


def _ragged(m):
    rng = np.random.default_rng(9)
    dna = lambda n: bytes(rng.choice(list(b"ACGT"), size=n).astype(np.uint8))
    site = b"AAATCGAACATGTGTTCGAGTA"
    return [dna(m - 1), dna(m), dna(m + 1), dna(3), b"N" * 700, dna(900).lower(), site, dna(100) + site + dna(31),
            dna(1023), dna(1024), dna(1025), dna(40000)]


def test_ragged_set_hits_and_background(pair):
    M, om = pair
    m = M.length
    seqs = _ragged(m)
    g = PackedGenome.from_sequences(seqs)
    assert g.total_windows(m) == sum(max(0, len(s) - m + 1) for s in seqs)
    # more than CGBK_INLINE_SEQS sequences: the set's tables live in the caller-owned device block
    assert g._table.numel() * 4 == 8 * int(_cabi.lib().cgbk_seqset_table_ints(len(seqs)))
    thr = 2.0                                       # low: plenty of hits in every sequence
    seq_i, pos, strand, score = M.scan_hits(g, thr)
    for k, s in enumerate(seqs):
        sel = seq_i == k
        if len(s) < m:
            assert not sel.any()
            continue
        ep, es, ev = c_oracle.scan_hits(om, s, thr)
        assert np.array_equal(pos[sel], ep) and np.array_equal(strand[sel], es) and np.array_equal(score[sel], ev), k
    b = np.concatenate([c_oracle.score_seq(om, s)[2] for s in seqs if len(s) >= m])
    for bins in (1, 256, 257, 4096):
        bg = M.background_stats(g, n_bins=bins)
        st = bg["stats"].cpu().numpy()
        assert st[_cabi.BG_N] == len(b) and int(bg["hist"].sum()) == len(b)
        assert abs(st[_cabi.BG_MEAN] - b.mean()) < 1e-6 and abs(st[_cabi.BG_STD] - b.std()) < 1e-6
    # a set with no window at all: n = 0, mean / std undefined
    g0 = PackedGenome.from_sequences([seqs[0], seqs[3]])
    st = M.background_stats(g0)["stats"].cpu().numpy()
    assert st[_cabi.BG_N] == 0 and np.isnan(st[_cabi.BG_MEAN])
    assert len(M.scan_hits(g0, thr)[1]) == 0


def test_slice_bounds_and_short_regions(pair, lexa):
    M, om = pair
    m = M.length
    seq = O.synthetic_genome(5000, 71, n_frac=0.002)
    g = PackedGenome.from_sequences([seq])
    M.fit_background(g)
    b = c_oracle.score_seq(om, seq)[2]
    O.build_bayesian_estimator(om, b)
    prior, alpha = lexa["prior_regulation_probability"], lexa["alpha"]
    # Python-slice semantics of Chromid.subsequence: negative start wraps once, end clamps
    starts = np.array([0, -300, 4800, 100, 4978, -22])
    ends = np.array([350, 5000, 6000, 100 + m, 5000, 5000])
    got = M.binding_probabilities(g, starts, ends, prior, alpha).cpu().numpy()
    for s, e, p in zip(starts, ends, got):
        sub = seq[int(s):int(e)]
        assert len(sub) >= m
        exp = c_oracle.binding_probability(om, sub, om._mu_m, om._std_m, om._mu_bg, om._std_bg, prior, alpha)
        assert abs(p - exp) < 1e-6, (s, e)
    # one base short of the motif: Biopython's calculate() returns an EMPTY score array
    # (numpy.empty(0)), so the likelihood ratio is exp(0) and the posterior is the prior ...
    got = M.binding_probabilities(g, [10, 0], [10 + m - 1, 350], prior, alpha).cpu().numpy()
    assert abs(got[0] - prior) < 1e-15
    assert M.binding_probability(seq[10:10 + m - 1].decode(), prior, alpha) == pytest.approx(prior, abs=1e-15)
    assert M.score_seq(seq[10:10 + m - 1].decode()) == []
    assert len(M.score_seq(seq[10:10 + m - 1].decode(), both=False)) == 0
    assert M.score_seqs([seq[:m - 1], seq[:m + 3]], True)[0] == [] and len(M.score_seqs([seq[:m - 1], seq[:m + 3]], True)[1]) == 4
    p2, _, _ = M.regulation_pipeline(seq, [10], [10 + m - 1], prior, alpha)
    assert abs(p2[0] - prior) < 1e-15
    with pytest.raises(ValueError):                                  # two bases short: numpy.empty(-1)
        M.binding_probabilities(g, [10], [10 + m - 2], prior, alpha)
    with pytest.raises(ValueError):
        M.score_seq(seq[10:10 + m - 2].decode())
    with pytest.raises(ValueError):
        M.regulation_pipeline(seq, [10], [10 + m - 2], prior, alpha)
    # no promoters at all: the pipeline still returns the background record
    post, stats, hist = M.regulation_pipeline(seq, [], [], prior, alpha)
    assert len(post) == 0 and stats[_cabi.BG_N] == len(b) and hist.sum() == len(b)
    assert abs(stats[_cabi.BG_MEAN] - b.mean()) < 1e-6


def test_pipeline_pageable_result_buffers(pair, lexa):
    """cgbk_regulation_pipeline writes results in place when the caller's buffers are pinned
    (the Python wrapper's case) and copies them back otherwise: same numbers either way."""
    import ctypes as C
    import torch
    M, om = pair
    seq = O.synthetic_genome(300000, 72, n_frac=0.001)
    starts = np.arange(0, 290000, 1000, dtype=np.int64)
    ends = starts + 350
    prior, alpha = lexa["prior_regulation_probability"], lexa["alpha"]
    post, stats, hist = M.regulation_pipeline(seq, starts, ends, prior, alpha)          # pinned buffers
    lib = _cabi.lib()
    n, n_bins = len(starts), 256
    seq_arr = np.frombuffer(seq, dtype=np.uint8)                                       # pageable input too
    scratch = torch.empty(int(lib.cgbk_pipeline_scratch_bytes(len(seq), n, M.length, n_bins)), dtype=torch.uint8,
                          device="cuda")
    stage = torch.empty(int(lib.cgbk_pipeline_staging_bytes(n)), dtype=torch.uint8).pin_memory()
    p2, s2, h2 = np.empty(n), np.empty(_cabi.BG_COUNT), np.empty(n_bins, dtype=np.uint64)   # separate, pageable
    lo, hi = M._score_bounds
    _cabi.check(lib.cgbk_regulation_pipeline(
        M._model(), seq_arr.ctypes.data, len(seq), starts.ctypes.data, ends.ctypes.data, n, float(M._mu_m),
        float(M._std_m), float(prior), float(alpha), float(lo), float(hi), n_bins, scratch.data_ptr(),
        scratch.numel(), stage.data_ptr(), stage.numel(), p2.ctypes.data, s2.ctypes.data, h2.ctypes.data,
        M._stream()))
    assert np.array_equal(h2, hist) and s2[_cabi.BG_N] == stats[_cabi.BG_N]
    assert abs(s2[_cabi.BG_MEAN] - stats[_cabi.BG_MEAN]) < 1e-12 and np.max(np.abs(p2 - post)) < 1e-12


def test_long_regions_beyond_the_marked_window_bitmask(pair, lexa):
    """Regions longer than 2 048 windows: high-scoring windows past the 64 x 32 positions the
    posterior kernel can merely mark are re-scored on the spot."""
    M, om = pair
    m = M.length
    seq = bytearray(O.synthetic_genome(30000, 73))
    site = b"AAATCGAACATGTGTTCGAGTA"
    for at in (150, 2500, 7000, 11990, 20000, 29900):
        seq[at:at + m] = site
    seq = bytes(seq)
    g = PackedGenome.from_sequences([seq])
    M.fit_background(g)
    b = c_oracle.score_seq(om, seq)[2]
    O.build_bayesian_estimator(om, b)
    prior, alpha = lexa["prior_regulation_probability"], lexa["alpha"]
    starts = np.array([0, 100, 5000, 0])
    ends = np.array([30000, 12100, 25000, 2100])
    for use_plane in (True, False):
        got = M.binding_probabilities(g, starts, ends, prior, alpha, use_plane=use_plane).cpu().numpy()
        for s, e, p in zip(starts, ends, got):
            exp = c_oracle.binding_probability(om, seq[int(s):int(e)], om._mu_m, om._std_m, om._mu_bg, om._std_bg,
                                               prior, alpha)
            assert abs(p - exp) < 1e-6, (use_plane, s, e, p, exp)


@pytest.mark.parametrize("n,lane", [(1_200_000, 32), (3_550_000, 64), (5_000_000, 96), (7_000_000, 128)])
def test_every_lane_segment_length_against_the_oracle(pair, lexa, n, lane):
    """The tiling is picked by a cost model (lane segments of 32 / 64 / 96 / 128 bases); the score
    plane, the histogram and the hit list must not care.  Sizes chosen to land on each."""
    M, om = pair
    seq = O.synthetic_genome(n, 80 + lane, n_frac=0.0002)
    g = PackedGenome.from_sequences([seq])
    M.fit_background(g, keep_scores=True)
    assert _cabi.last_launch_info()["lane_bases"] == lane
    b = c_oracle.score_seq(om, seq)[2]
    O.build_bayesian_estimator(om, b)
    assert abs(M._mu_bg - om._mu_bg) < 1e-6 and abs(M._std_bg - om._std_bg) < 1e-6
    rng = np.random.default_rng(lane)
    starts = np.sort(rng.integers(0, n - 400, size=300))
    ends = starts + rng.integers(M.length, 400, size=300)
    prior, alpha = lexa["prior_regulation_probability"], lexa["alpha"]
    got = M.binding_probabilities(g, starts, ends, prior, alpha).cpu().numpy()
    exp = np.array([c_oracle.binding_probability(om, seq[int(s):int(e)], om._mu_m, om._std_m, om._mu_bg,
                                                 om._std_bg, prior, alpha) for s, e in zip(starts, ends)])
    assert np.max(np.abs(got - exp)) < 1e-6
    _, pos, strand, score = M.scan_hits(g)
    ep, es, ev = c_oracle.scan_hits(om, seq, om.threshold())
    assert np.array_equal(pos, ep) and np.array_equal(strand, es) and np.array_equal(score, ev)


def test_packed_genome_file_round_trip(pair, tmp_path):
    """PackedGenome.save / load: the on-disk packed form scans exactly like the freshly packed one
    (hits incl. the ambiguity repair, background record)."""
    M, om = pair
    seqs = [O.synthetic_genome(300000, 91, n_frac=0.001), b"acgtnACGT" * 50, O.synthetic_genome(70001, 92)]
    g = PackedGenome.from_sequences(seqs)
    path = str(tmp_path / "genome.npz")
    g.save(path, names=["chr", "tiny", "plasmid"])
    g2, names = PackedGenome.load(path)
    assert names == ["chr", "tiny", "plasmid"] and g2.seq_len.tolist() == g.seq_len.tolist()
    a, b = M.scan_hits(g, 3.0), M.scan_hits(g2, 3.0)
    assert all(np.array_equal(x, y) for x, y in zip(a, b)) and len(a[1]) > 100
    sa, sb = M.background_stats(g), M.background_stats(g2)
    assert torch_equal(sa["hist"], sb["hist"]) and torch_equal(sa["stats"][:1], sb["stats"][:1])
    with pytest.raises(ValueError):
        PackedGenome([10, 20]).save(path)                       # nothing packed yet


def torch_equal(a, b):
    import torch
    return bool(torch.equal(a, b))

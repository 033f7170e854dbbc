"""BASELINE.json configs[2..4] at reduced size, against the oracle (GPU parity tests).

configs[2]  many genomes x 3 derived PSWMs, genomes sharded over ranks (no collective)
configs[3]  one long sequence, chunk-sharded with an (m-1)-base halo, background record summed
configs[4]  many JASPAR-scale motifs (8-30 bp) over one sequence
The multi-rank plumbing itself is covered by tests/test_dist_cpu.py (gloo) and
tests/test_gpu_dist.py (NCCL, 2 GPUs); here the ranks are played one after the other on
one GPU so that the data path (chunking, accumulation, per-genome models) is checked.
"""
import numpy as np
import pytest
import torch

from cgb_b200 import GeneTable, PackedGenome, PSSMModel, SiteCollection, _cabi
from cgb_b200 import dist as cdist
from cgb_b200 import calculate_regulation_probabilities
from oracle import c_oracle, np_oracle as O

pytestmark = pytest.mark.gpu


def _collections(lexa):
    return [SiteCollection(mo["sites"], None, mo["name"]) for mo in lexa["motifs"]]


def test_config3_genomes_x_three_pswms(lexa, lexa_weights):
    site_lists = [mo["sites"] for mo in lexa["motifs"]]
    sizes = [180000, 150001, 210000, 120000, 199999, 170000]
    parts = cdist.assign_genomes(sizes, 2)
    assert sorted(i for p in parts for i in p) == list(range(len(sizes)))
    prior, alpha = lexa["prior_regulation_probability"], lexa["alpha"]
    for rank, mine in enumerate(parts):               # "ranks", one after the other
        for gi in mine:
            seq = O.synthetic_genome(sizes[gi], 100 + gi, n_frac=0.0005)
            genes = O.synthetic_genes(sizes[gi], 60)
            gt = GeneTable([x.start for x in genes], [x.end for x in genes], [x.strand for x in genes], sizes[gi])
            g = PackedGenome.from_sequences([seq])
            rng = np.random.default_rng(gi)
            for w in (lexa_weights["equal"], lexa_weights["site_count"], rng.dirichlet([1.0] * 6).tolist()):
                M = PSSMModel(_collections(lexa), w)
                om = O.OracleModel(site_lists, w)
                M.fit_background(g)
                _, _, b = c_oracle.score_seq(om, seq)
                O.build_bayesian_estimator(om, b)
                assert abs(M._mu_bg - om._mu_bg) < 1e-6 and abs(M._std_bg - om._std_bg) < 1e-6
                post = calculate_regulation_probabilities(M, g, gt, prior, alpha).cpu().numpy()
                exp = []
                for gene in genes:
                    s, e = O.promoter_region_location(gene, sizes[gi], 300, 50)
                    exp.append(c_oracle.binding_probability(om, seq[s:e], om._mu_m, om._std_m, om._mu_bg,
                                                            om._std_bg, prior, alpha))
                assert np.max(np.abs(post - np.array(exp))) < 1e-6
                _, pos, strand, score = M.scan_hits(g)
                ep, es, ev = c_oracle.scan_hits(om, seq, om.threshold())
                assert np.array_equal(pos, ep) and np.array_equal(strand, es) and np.array_equal(score, ev)


def test_config4_chunk_sharded_sequence(lexa, lexa_weights):
    site_lists = [mo["sites"] for mo in lexa["motifs"]]
    M = PSSMModel(_collections(lexa), lexa_weights["site_count"])
    om = O.OracleModel(site_lists, lexa_weights["site_count"])
    n, world = 3_000_017, 8
    seq = O.synthetic_genome(n, 4, n_frac=0.0002)
    starts, ends, counts = cdist.chunk_bounds(n, M.length, world)
    bg = None
    hits = []
    for r in range(world):                            # "ranks", one after the other
        g = PackedGenome.from_sequences([seq[starts[r]:ends[r]]])
        assert g.total_windows(M.length) == counts[r]
        bg = M.background_stats(g, n_bins=512, accumulate_into=bg)     # plays the all-reduce (sum)
        _, pos, strand, score = M.scan_hits(g, M.threshold() - 2.0)
        hits.append((pos + starts[r], strand, score))
    whole = M.background_stats(PackedGenome.from_sequences([seq]), n_bins=512)
    a, w = bg["stats"].cpu().numpy(), whole["stats"].cpu().numpy()
    _, _, b = c_oracle.score_seq(om, seq)
    assert a[_cabi.BG_N] == w[_cabi.BG_N] == len(b)
    # integer moments make the tiled scan independent of the chunking; only the windows of
    # "dirty" tiles (exact float64 path) change hands between the two tilings
    assert abs(a[_cabi.BG_MEAN] - w[_cabi.BG_MEAN]) < 1e-9 and abs(a[_cabi.BG_STD] - w[_cabi.BG_STD]) < 1e-9
    assert abs(a[_cabi.BG_MEAN] - b.mean()) < 1e-6 and abs(a[_cabi.BG_STD] - b.std()) < 1e-6
    assert int((bg["hist"] - whole["hist"]).abs().sum().item()) <= 40
    assert int(bg["hist"].sum().item()) == len(b)
    pos = np.concatenate([h[0] for h in hits])
    strand = np.concatenate([h[1] for h in hits])
    score = np.concatenate([h[2] for h in hits])
    order = np.lexsort((-strand, pos))
    ep, es, ev = c_oracle.scan_hits(om, seq, om.threshold() - 2.0)
    assert np.array_equal(pos[order], ep) and np.array_equal(strand[order], es) and np.array_equal(score[order], ev)


def test_config5_many_motifs_one_sequence():
    rng = np.random.default_rng(5)
    seq = O.synthetic_genome(400000, 5, n_frac=0.0003)
    g = PackedGenome.from_sequences([seq])
    total_hits = 0
    models, singles = [], []
    for k in range(24):
        m = int(rng.integers(8, 31))
        # JASPAR-like: 20 sites sampled from a Dirichlet(0.5)-column PWM, pseudocount 1
        cols = rng.dirichlet([0.5] * 4, size=m)
        sites = ["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m)) for _ in range(20)]
        M = PSSMModel([SiteCollection(sites)], [1.0])
        om = O.OracleModel([sites], [1.0])
        assert M.threshold() == om.threshold() and M.IC == om.IC
        _, pos, strand, score = M.scan_hits(g)
        ep, es, ev = c_oracle.scan_hits(om, seq, om.threshold())
        assert np.array_equal(pos, ep) and np.array_equal(strand, es) and np.array_equal(score, ev), (k, m)
        total_hits += len(pos)
        models.append(M)
        singles.append((pos, strand, score))
        st = M.background_stats(g)["stats"].cpu().numpy()
        _, _, b = c_oracle.score_seq(om, seq)
        assert abs(st[_cabi.BG_MEAN] - b.mean()) < 1e-6 and abs(st[_cabi.BG_STD] - b.std()) < 1e-6, (k, m)
    assert total_hits > 100
    # all motifs enqueued back to back, collected afterwards
    from cgb_b200 import scan_hits_many
    for (pos, strand, score), res in zip(singles, scan_hits_many(models, g)):
        assert np.array_equal(res[1], pos) and np.array_equal(res[2], strand) and np.array_equal(res[3], score)


def test_config2_full_size(lexa, lexa_weights):
    """BASELINE configs[1] at its full size: 5 Mbp genome, 3 000 promoters, against the C oracle."""
    site_lists = [mo["sites"] for mo in lexa["motifs"]]
    M = PSSMModel(_collections(lexa), lexa_weights["site_count"])
    om = O.OracleModel(site_lists, lexa_weights["site_count"])
    L, n_genes = 5_000_000, 3000
    seq = O.synthetic_genome(L, 2)
    step = L // n_genes
    start = np.arange(n_genes, dtype=np.int64) * step + step // 3
    end = np.minimum(L, start + min(1000, step // 2))
    strand = np.where(np.arange(n_genes) % 2 == 0, 1, -1)
    gt = GeneTable(start, end, strand, L)
    ps, pe = gt.promoter_region_location(300, 50)
    prior, alpha = lexa["prior_regulation_probability"], lexa["alpha"]
    post, stats, hist = M.regulation_pipeline(seq, ps, pe, prior, alpha, n_bins=1024)
    b = c_oracle.score_seq(om, seq)[2]
    O.build_bayesian_estimator(om, b)
    assert stats[_cabi.BG_N] == len(b) == L - 21 and hist.sum() == len(b)
    assert abs(stats[_cabi.BG_MEAN] - om._mu_bg) < 1e-6 and abs(stats[_cabi.BG_STD] - om._std_bg) < 1e-6
    exp = np.array([c_oracle.binding_probability(om, seq[int(s):int(e)], om._mu_m, om._std_m, om._mu_bg,
                                                 om._std_bg, prior, alpha) for s, e in zip(ps, pe)])
    assert np.max(np.abs(post - exp)) < 1e-6


@pytest.mark.parametrize("n,chunks", [(1_600_000, 3), (3_000_001, 4), (17_000_000, None)])
def test_pipeline_chunked_upload_matches_resident_path(lexa, lexa_weights, n, chunks, monkeypatch):
    """cgbk_regulation_pipeline uploads large sequences in chunks (4 from 16 MB on; forced here
    for the small cases through the CGBK_PIPE_CHUNKS tuning knob) and scans each chunk as it
    arrives: same histogram / count as the resident whole-genome scan, also with non-ACGT runs
    sitting on the chunk cuts (those tiles go to the exact kernel)."""
    if chunks:
        monkeypatch.setenv("CGBK_PIPE_CHUNKS", str(chunks))
    M = PSSMModel(_collections(lexa), lexa_weights["site_count"])
    seq = np.frombuffer(O.synthetic_genome(n, 77), dtype=np.uint8).copy()
    fracs = (0.5, 0.8) if chunks == 3 else (0.35, 0.65, 0.85)
    for f in fracs:
        cut = int(f * n) // 512 * 512
        seq[cut - 7:cut + 9] = ord("N")
        seq[cut + 3000:cut + 3002] = ord("n")
    seq[100:140] = np.frombuffer(b"acgtacgtacgtacgtacgtacgtacgtacgtacgtacgt", dtype=np.uint8)   # lower case: repaired
    n_genes = 500
    step = n // n_genes
    start = np.arange(n_genes, dtype=np.int64) * step + step // 3
    gt = GeneTable(start, start + 900, np.where(np.arange(n_genes) % 2 == 0, 1, -1), n)
    ps, pe = gt.promoter_region_location(300, 50)
    prior, alpha = lexa["prior_regulation_probability"], lexa["alpha"]
    post, stats, hist = M.regulation_pipeline(seq, ps, pe, prior, alpha, n_bins=256)
    M2 = PSSMModel(_collections(lexa), lexa_weights["site_count"])
    g = PackedGenome.from_sequences([seq])
    bg = M2.background_stats(g, n_bins=256, keep_scores=True)
    M2.fit_background(bg=bg)
    st2 = bg["stats"].cpu().numpy()
    assert np.array_equal(hist, bg["hist"].cpu().numpy().astype(np.uint64))
    assert stats[_cabi.BG_N] == st2[_cabi.BG_N] == n - M.length + 1
    assert abs(stats[_cabi.BG_MEAN] - st2[_cabi.BG_MEAN]) < 1e-12
    assert abs(stats[_cabi.BG_STD] - st2[_cabi.BG_STD]) < 1e-12
    post2 = M2.binding_probabilities(g, ps, pe, prior, alpha).cpu().numpy()
    assert np.max(np.abs(post - post2)) < 1e-12


def test_large_sequence_tiling_invariance(lexa, lexa_weights):
    """Size-independent properties at 200 Mbp (no oracle at this size): the hit list and the
    background record do not depend on how the sequence is cut into chunks / tiles, every window
    is counted exactly once, and hits are sorted and unique."""
    M = PSSMModel(_collections(lexa), lexa_weights["site_count"])
    n, world = 200_000_000, 8
    gen = torch.Generator(device="cuda")
    gen.manual_seed(44)
    lut = torch.tensor(list(b"ACGT"), dtype=torch.uint8, device="cuda")
    ascii_all = lut[torch.randint(0, 4, (n,), device="cuda", generator=gen)]
    whole = PackedGenome([n])
    whole.pack_into(0, ascii_all)
    _, pos, strand, score = M.scan_hits(whole)
    bg_whole = M.background_stats(whole, n_bins=256)
    key = pos * 2 + (strand < 0)
    assert (np.diff(key) > 0).all()                                   # sorted, no duplicates
    assert (score.astype(np.float64) >= M.threshold()).all()
    starts, ends, counts = cdist.chunk_bounds(n, M.length, world)
    bg = None
    parts = []
    for r in range(world):
        g = PackedGenome([int(ends[r] - starts[r])])
        g.pack_into(0, ascii_all[int(starts[r]):int(ends[r])].clone())
        bg = M.background_stats(g, n_bins=256, accumulate_into=bg)
        _, p, s, v = M.scan_hits(g)
        parts.append((p + starts[r], s, v))
    assert np.array_equal(np.concatenate([x[0] for x in parts]), pos)
    assert np.array_equal(np.concatenate([x[1] for x in parts]), strand)
    assert np.array_equal(np.concatenate([x[2] for x in parts]), score)
    a, w = bg["stats"].cpu().numpy(), bg_whole["stats"].cpu().numpy()
    assert a[_cabi.BG_N] == w[_cabi.BG_N] == n - M.length + 1
    # integer moments inside a scan; the sums of the 8 chunk calls are then added in float64
    assert abs(a[_cabi.BG_SUM] - w[_cabi.BG_SUM]) <= 1e-14 * abs(w[_cabi.BG_SUM])
    assert abs(a[_cabi.BG_MEAN] - w[_cabi.BG_MEAN]) < 1e-12 and abs(a[_cabi.BG_STD] - w[_cabi.BG_STD]) < 1e-12
    assert torch.equal(bg["hist"], bg_whole["hist"])
    assert int(bg["hist"].sum().item()) == n - M.length + 1

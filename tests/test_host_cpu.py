"""CPU-side checks of the product: host float64 motif algebra against the golden
vectors, the reference-facing API surface, coordinate rules, and that libcgbk.so
loads and exports every symbol include/cgbk.h declares (no GPU compute here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from cgb_b200 import GeneTable, PSSMModel, SiteCollection, TFBindingModel, _cabi, python_slice_bounds
from oracle import np_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _collections(lexa):
    return [SiteCollection(mo["sites"], None, mo["name"]) for mo in lexa["motifs"]]


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "cgbk.h")).read()
    declared = set(re.findall(r"\b(cgbk_[a-z0-9_]+)\s*\(", header))
    assert declared == set(_cabi.SYMBOLS), declared ^ set(_cabi.SYMBOLS)
    lib = C.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _cabi.lib().cgbk_version() == 211


def test_site_collection_pwm(golden, lexa):
    for c, g in zip(_collections(lexa), golden["collections"]):
        assert c.site_count == g["n_sites"] and c.length == 22
        for l in "ACGT":
            assert list(c.pwm[l]) == g["pwm"][l]
        # the golden IC / threshold are those of PSSMModel([c], [1.0]) (re-normalised mix)
        M = PSSMModel([c], [1.0])
        assert M.IC == g["IC"] and M.threshold() == g["patser_threshold"]
        assert abs(c.IC - g["IC"]) < 1e-12
        assert c.sites[0] == g["first_site"]


@pytest.mark.parametrize("wname", ["equal", "site_count"])
def test_pssm_model_host_math(golden, lexa, lexa_weights, wname):
    rec = golden["models"][wname]
    M = PSSMModel(_collections(lexa), lexa_weights[wname])
    assert isinstance(M, TFBindingModel)
    assert M.length == 22 and M.alphabet == "ACGT" and M.background["A"] == 0.25
    for l in "ACGT":
        assert list(M.pwm[l]) == rec["pwm"][l]
        assert list(M.pssm[l]) == rec["pssm"][l]
        assert list(M.rev_comp_pssm[l]) == rec["rev_comp_pssm"][l]
    assert M.IC == rec["IC"]
    assert M.threshold() == rec["patser_threshold"] == M.patser_threshold
    with pytest.raises(ValueError):
        M.threshold("other")
    assert len(M.sites) == 164 and M.site_collections[0].name == "LexA_Mtu"
    with pytest.raises(AttributeError):
        M.bayesian_estimator                       # never assigned in the reference either
    # host restatement of the ambiguity rule (pssm_model.py:88-101)
    om = O.OracleModel([mo["sites"] for mo in lexa["motifs"]], lexa_weights[wname])
    site = "gaatcaaacNTGTGTTCGACAG"
    assert M._calculate(M.pssm, site) == O._calculate_ambiguous(om.fwd, site.encode())
    assert M._calculate(M.rev_comp_pssm, site) == O._calculate_ambiguous(om.rc, site.encode())


def test_patser_threshold_cabi_matches_oracle(lexa):
    rng = np.random.default_rng(5)
    for m in (8, 13, 30):
        pwm = rng.dirichlet([0.5] * 4, size=m) * 0.9 + 0.025
        lo = np.log2(pwm / 0.25)
        om = O.OracleModel.from_logodds(lo)
        flat = (C.c_double * (4 * m))(*lo.ravel().tolist())
        thr, ic, mn, mx = C.c_double(), C.c_double(), C.c_double(), C.c_double()
        _cabi.check(_cabi.lib().cgbk_patser_threshold(flat, m, 1000, C.byref(thr), C.byref(ic),
                                                      C.byref(mn), C.byref(mx)))
        assert ic.value == O.pssm_mean(om.pssm)
        assert mn.value == O.pssm_min(om.pssm) and mx.value == O.pssm_max(om.pssm)
        assert abs(thr.value - O.patser_threshold(om.pssm)) < 1e-9


def test_errors_map_to_reference_exception_types():
    lo = (C.c_double * 8)(*([0.1] * 8))
    with pytest.raises(ValueError):
        _cabi.check(_cabi.lib().cgbk_patser_threshold(lo, 0, 1000, None, None, None, None))
    bad = (C.c_double * 4)(0.0, float("-inf"), 0.0, 0.0)
    with pytest.raises(ValueError):
        _cabi.check(_cabi.lib().cgbk_patser_threshold(bad, 1, 1000, None, None, None, None))
    a = SiteCollection(["ACGT", "ACGA"])
    b = SiteCollection(["ACGTT"])
    with pytest.raises(AssertionError):
        PSSMModel([a, b], [0.5, 0.5])               # pssm_model.py:164
    with pytest.raises(KeyError):
        SiteCollection(["ACGN"])                     # Biopython counts only alphabet letters


def test_gene_table_matches_oracle_rules():
    rng = np.random.default_rng(7)
    L = 50000
    starts = np.sort(rng.choice(np.arange(10, L - 2000), size=40, replace=False))
    ends = np.minimum(starts + rng.integers(60, 1500, size=40), L)
    strand = rng.choice([1, -1], size=40)
    # edge cases: gene at the very start / end of the replicon
    starts[0], ends[0], strand[0] = 5, 300, -1
    ends[-1], strand[-1] = L - 3, -1
    gt = GeneTable(starts, ends, strand, L)
    genes = [O.OGene(int(s), int(e), int(t)) for s, e, t in zip(starts, ends, strand)]
    for up, down in ((None, 50), (300, 50), (0, 50), (1000, 0), (250, 100)):
        s, e = gt.upstream_noncoding_region_location(up, down)
        for i in range(len(genes)):
            assert (int(s[i]), int(e[i])) == O.upstream_noncoding_region_location(genes, i, L, up, down)
    for up, down in ((300, 50), (0, 0), (1000, 1000)):
        s, e = gt.promoter_region_location(up, down)
        for i in range(len(genes)):
            assert (int(s[i]), int(e[i])) == O.promoter_region_location(genes[i], L, up, down)


def test_python_slice_bounds():
    seq = "ACGT" * 25
    L = len(seq)
    cases = [(0, 10), (-10, 5), (-10, 200), (90, 300), (-300, 20), (50, 40), (120, 130), (-5, -1)]
    s, e = python_slice_bounds([c[0] for c in cases], [c[1] for c in cases], L)
    for (a, b), si, ei in zip(cases, s, e):
        assert seq[a:b] == seq[si:ei], (a, b)


def test_scoring_without_gpu_fails_loudly():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    M = PSSMModel([SiteCollection(["ACGTAC", "ACGTAA"])], [1.0])
    with pytest.raises(RuntimeError):
        M.score_seq("ACGTACGTAC")


def test_identified_sites_rows_match_reference_format(tmp_path):
    """the identified_sites CSV rows (genome.py:457-494) from Site tuples."""
    import csv

    from cgb_b200 import Site
    from cgb_b200.site_search import (CSV_HEADER, identified_sites_rows, relative_distance_to_start,
                                      write_identified_sites_csv)
    rng = np.random.default_rng(11)
    L = 5000
    seq = O.synthetic_genome(L, 12).decode()
    genes = O.synthetic_genes(L, 6)
    ann = [("tag%d" % i, "WP_%05d.1" % i if i % 2 else "", "product %d" % i) for i in range(len(genes))]
    gt = GeneTable([g.start for g in genes], [g.end for g in genes], [g.strand for g in genes], L,
                   locus_tag=[a[0] for a in ann], protein_id=[a[1] for a in ann], product=[a[2] for a in ann])
    sites, osites = [], []
    for _ in range(40):
        gi = int(rng.integers(0, len(genes)))
        start = int(rng.integers(0, L - 22))
        strand = int(rng.choice([1, -1]))
        score = np.float32(rng.uniform(9, 25))
        sites.append(Site(0, start, start + 22, strand, score, gi))
        osites.append(O.OSite(start, start + 22, strand, score, gi))
    for up in (300, 50):
        got = identified_sites_rows(sites, [seq], [gt], up, accessions=["NC_TEST.1"])
        exp = O.identified_sites_rows(osites, seq, genes, up, "NC_TEST.1", ann)
        assert got == exp
    for g in genes:      # every branch of gene.py:290-316
        for s, e in ((g.start - 40, g.start - 18), (g.start, g.start + 22), (g.start + 5, g.start + 27),
                     (g.end - 22, g.end), (g.end, g.end + 22), (g.end - 5, g.end + 17)):
            assert relative_distance_to_start(g.start, g.end, g.strand, s, e) == O.relative_distance_to_start(g, s, e)
    path = tmp_path / "sites.csv"
    write_identified_sites_csv(str(path), sites, [seq], [gt], 300, accessions=["NC_TEST.1"])
    rows = list(csv.reader(open(path)))
    assert rows[0] == CSV_HEADER and len(rows) == 41
    assert rows[1][1] == got[0][1] and rows[1][7] in ("operator", "intergenic")


@pytest.mark.parametrize("wname", ["equal", "site_count"])
def test_score_self_on_the_host_matches_verbatim_reference(golden, lexa, lexa_weights, wname):
    """PSSMModel.score_self goes through cgbk_score_sites_host (float64 on the host, no GPU): the
    164 LexA self scores against the verbatim reference run, and mu_m / std_m from them."""
    rec = golden["models"][wname]
    M = PSSMModel(_collections(lexa), lexa_weights[wname])
    got = M.score_self()
    assert len(got) == len(rec["score_self"]) and all(len(x) == 1 for x in got)
    # the fixture's soft-max ran with numpy 2's float32 pow; the pinned numpy 1.x (what the
    # library follows) differs by < 5e-7 per score
    assert np.max(np.abs(np.array([x[0] for x in got]) - np.array(rec["score_self"]))) < 5e-7
    om = O.OracleModel([mo["sites"] for mo in lexa["motifs"]], lexa_weights[wname])
    exp = O.score_self(om)                                    # legacy64 mode: the same arithmetic
    assert np.max(np.abs(np.array([x[0] for x in got]) - np.array([x[0] for x in exp]))) < 1e-13
    # non-ACGT and lower-case bytes: _calculate's repair (pssm_model.py:88-101), float64 kept
    m = M.length
    site = M.sites[0]
    for edited in (site[:4] + "N" + site[5:], site.lower(), site[:2] + "n" + site[3:8].lower() + site[8:]):
        M2 = PSSMModel(_collections(lexa), lexa_weights[wname])
        M2.__dict__["sites"] = [edited]
        assert abs(M2.score_self()[0][0] - O.score_seq(om, edited, True)[0][0]) < 1e-13
    assert m == 22


def test_split_sites_rebuild_records_on_the_host():
    """SplitSites (the 6-byte transfer form of a site list) against a numpy statement of the split:
    positions, strands and the 8-byte records come back exactly, empty blocks and an empty list too."""
    from cgb_b200 import motif_batch as mb
    rng = np.random.default_rng(5)
    cap = 5 * 4096 + 512
    for n in (0, 1, 37, 4000):
        pos = np.sort(rng.integers(0, cap, size=n)).astype(np.int64)
        if n > 30:
            pos[10:20] = pos[10]                       # both strands at one position, a block with many sites
            pos = np.sort(pos)
        minus = rng.integers(0, 2, size=n).astype(np.uint64)
        score = rng.normal(0, 5, size=n).astype(np.float32)
        rec = (pos.astype(np.uint64) << np.uint64(33)) | (minus << np.uint64(32)) | score.view(np.uint32).astype(np.uint64)
        nb = (cap + 4095) // 4096
        local = (((pos & 4095) << 1) | minus.astype(np.int64)).astype(np.uint16)
        block_start = np.searchsorted(pos, np.arange(nb + 1, dtype=np.int64) << 12, side="left").astype(np.int32)
        s = mb.SplitSites(local, score, block_start)
        assert len(s) == n and np.array_equal(s.records(), rec) and np.array_equal(s.positions(), pos)
        assert np.array_equal(s.strands(), 1 - 2 * minus.astype(np.int32))
        gp, st, sc = mb.unpack_hits(s.records())
        assert np.array_equal(gp, pos) and np.array_equal(sc, score)


def test_interleave_order_tracks_the_batch_ratio():
    """motif_batch.interleave_order: a permutation, deterministic, and along the processing order the
    sites sent stay within one heavy model of (overall ratio x device time spent)."""
    from cgb_b200 import motif_batch as mb
    rng = np.random.default_rng(2)
    for n in (0, 1, 2, 3, 17, 200):
        cost = rng.uniform(1.0, 2.5, size=n)
        sites = np.where(rng.random(n) < 0.2, rng.uniform(5e7, 3e8, size=n), rng.uniform(1e3, 2e6, size=n))
        order = mb.interleave_order(cost.tolist(), sites.tolist())
        assert sorted(order) == list(range(n)) and order == mb.interleave_order(list(cost), list(sites))
        if n >= 17:
            ratio = sites.sum() / cost.sum()
            gap = np.abs(np.cumsum(sites[order]) - ratio * np.cumsum(cost[order]))
            assert np.all(gap[:-1] <= sites.max() + ratio * cost.max())
            assert sites[order[0]] == sites.max()          # a PCIe-heavy model first: the copies start early
    assert mb.interleave_order([1.0, 1.0, 1.0], [0, 0, 0]) == [0, 1, 2]


def test_score_self_with_foreign_collection_types():
    """score_self takes the collections' cached site bytes only when every collection is a
    cgb_b200 SiteCollection; PSSMModel.from_pwm (a bare PWM with optional sites) and duck-typed
    collections go through ``.sites`` and give the same numbers."""
    vals = {"A": [0.7, 0.1, 0.1, 0.25], "C": [0.1, 0.7, 0.1, 0.25], "G": [0.1, 0.1, 0.7, 0.25], "T": [0.1, 0.1, 0.1, 0.25]}
    sites = ["ACGT", "ACGA", "TCGT", "ACGT"]
    a = PSSMModel.from_pwm(vals, sites=sites).score_self()
    assert len(a) == 4 and a[0] == a[3] and a[0][0] > a[2][0]
    assert PSSMModel.from_pwm(vals).score_self() == []

    class Duck(object):                                    # what a caller's own collection class offers
        def __init__(self, sites):
            self._c = SiteCollection(sites)
            self.pwm, self.sites, self.site_count, self.length = self._c.pwm, self._c.sites, len(sites), 4
    mine = PSSMModel([SiteCollection(sites)], [1.0]).score_self()
    duck = PSSMModel([Duck(sites)], [1.0]).score_self()
    assert mine == duck and len(mine) == 4


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside ours) on a tiny sample: one
    JSON line with the contract's keys, our metric / unit, impl = reference, a cpu_baseline object
    describing the run and an e2e object repeating the value with zero transfer bytes."""
    import json
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--steps", "2",
                        "--warmup", "1", "--ref-sample-bp", "2000", "--motifs", "5"],
                       capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-800:]
    lines = [l for l in r.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "bp*motif/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2 and d["n_gpus"] == 1 and d["gpu_launches"] == 0
    assert d["cpu_baseline"]["kind"] in ("port", "reference") and d["cpu_baseline"]["cores"] >= 1
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"]
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    assert "workload" in d["config"] and d["metric"].startswith("PSWM scan")

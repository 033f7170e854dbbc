"""Genome-side rows of SURVEY.md §8 against the reference run verbatim (CPU, no GPU needed).

tests/golden/genome_vectors.json holds what the unmodified cgb/genome.py, chromid.py, gene.py and
operon.py produce on a small synthetic annotated genome (tests/golden/make_genome_golden.py).
Checked here: the GenBank parser + gene-selection rules, the coordinate rules (product and
oracle), the oracle's identify_sites / report rows / posterior, and operon prediction + CSVs.
"""
import csv
import io
import json
import os

import numpy as np
import pytest

from cgb_b200 import GeneTable, genbank, operons, python_slice_bounds
from oracle import np_oracle as O
from genbank_writer import genbank_text

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def gv():
    with open(os.path.join(HERE, "golden", "genome_vectors.json")) as f:
        return json.load(f)


def _table(gold, k=0):
    rows = gold["rows"]
    col = lambda name: [r[name] for r in rows]
    return GeneTable(col("start"), col("end"), col("strand"), gold["length"], seq_index=k,
                     locus_tag=col("locus_tag"), protein_id=col("protein_id"), product=col("product"),
                     index=col("index"), name=col("name"), product_type=col("product_type"))


def _csv_text(rows):
    buf = io.StringIO(newline="")
    w = csv.writer(buf)
    for r in rows:
        w.writerow(r)
    return buf.getvalue()


def test_genbank_parse_matches_reference_gene_table(gv):
    """GenBank text -> sequence + GeneTable == what Chromid.genes derives from the same features."""
    for rep, gold in zip(gv["replicons"], gv["genes"]):
        rec = genbank.read(genbank_text(rep))
        assert rec.id == rep["accession"] and rec.description == rep["description"]
        assert rec.sequence.decode() == rep["sequence"].upper()       # Biopython upper-cases ORIGIN
        assert len(rec.features) == len(rep["features"])
        gt = rec.gene_table()
        rows = gold["rows"]
        assert len(gt) == len(rows)
        for name in ("index", "start", "end", "strand"):
            assert getattr(gt, name).tolist() == [r[name] for r in rows], name
        for name in ("locus_tag", "name", "product_type", "product", "protein_id"):
            assert list(getattr(gt, name)) == [r[name] for r in rows], name
        # the oracle's restatement of chromid.py:99-122 on the raw feature dicts
        orows = O.genes_from_features(rep["features"])
        assert [(r["index"], r["start"], r["end"], r["strand"], r["locus_tag"], r["product_type"], r["product"],
                 r["protein_id"]) for r in rows] == orows


def test_genbank_location_forms():
    f = genbank.parse_location
    assert f("123..456") == (122, 456, 1, False)
    assert f("complement(123..456)") == (122, 456, -1, False)
    assert f("<123..>456") == (122, 456, 1, False)
    assert f("77") == (76, 77, 1, False)
    assert f("join(1..10,20..30)")[3] and f("complement(join(1..10,\n 20..30))")[2:] == (-1, True)
    assert f("order(5..9,12..14)")[3]
    assert f("J00194.1:100..202") == (99, 202, 1, False)
    with pytest.raises(ValueError):
        genbank.read("no records here")
    with pytest.raises(ValueError):
        rep = {"accession": "X.1", "description": "d", "sequence": "ACGT", "features": []}
        genbank.read(genbank_text(rep) + genbank_text(rep))


def test_coordinate_rules_match_reference(gv):
    up, dw = gv["promoter_up_distance"], gv["promoter_dw_distance"]
    for k, (rep, gold) in enumerate(zip(gv["replicons"], gv["genes"])):
        gt = _table(gold, k)
        rows = gold["rows"]
        tags = [r["locus_tag"] for r in rows]
        ug = gt.upstream_gene()
        assert [tags[i] if i >= 0 else None for i in ug] == [r["upstream_gene"] for r in rows]
        s, e = gt.upstream_noncoding_region_location(None, dw)
        assert [[int(a), int(b)] for a, b in zip(s, e)] == [r["upstream_noncoding_none"] for r in rows]
        s, e = gt.upstream_noncoding_region_location(up, dw)
        assert [[int(a), int(b)] for a, b in zip(s, e)] == [r["upstream_noncoding_up"] for r in rows]
        s, e = gt.promoter_region_location(up, dw)
        s, e = python_slice_bounds(s, e, gold["length"])
        assert [rep["sequence"][int(a):int(b)] for a, b in zip(s, e)] == [r["promoter_region"] for r in rows]
        # the oracle's scalar rules
        genes = [O.OGene(r["start"], r["end"], r["strand"]) for r in rows]
        index = [r["index"] for r in rows]
        for i, r in enumerate(rows):
            assert list(O.upstream_noncoding_region_location(genes, i, gold["length"], None, dw, index)) == \
                r["upstream_noncoding_none"]
            assert list(O.upstream_noncoding_region_location(genes, i, gold["length"], up, dw, index)) == \
                r["upstream_noncoding_up"]
            a, b = O.promoter_region_location(genes[i], gold["length"], up, dw)
            assert O.subsequence(rep["sequence"], a, b) == r["promoter_region"]


def test_upstream_gene_index_error_like_reference():
    # a forward gene whose feature number points past the shortened list: genes[_index - 1] raises
    gt = GeneTable([100, 900], [500, 1300], [1, 1], 5000, index=[0, 4])
    with pytest.raises(IndexError):
        gt.upstream_noncoding_region_location(None, 50)
    s, e = gt.upstream_noncoding_region_location(300, 50)          # never looks the neighbour up
    assert s.tolist() == [0, 600] and e.tolist() == [150, 950]


def _oracle_model(gv, lexa):
    om = O.OracleModel([mo["sites"] for mo in lexa["motifs"]], gv["weights"])
    est = gv["estimator"]
    om._mu_m, om._std_m, om._mu_bg, om._std_bg = est["mu_m"], est["std_m"], est["mu_bg"], est["std_bg"]
    return om


def test_oracle_identify_sites_and_report_match_reference(gv, lexa):
    """Pins the oracle's hit loop and report rows (genome.py:394-494) to the verbatim run:
    same sites in the same order, CSV text byte for byte."""
    om = _oracle_model(gv, lexa)
    assert om.threshold() == gv["estimator"]["threshold"]
    up, dw = gv["promoter_up_distance"], gv["promoter_dw_distance"]
    for scen, use_up in (("scan_to_upstream_gene", None), ("scan_up_distance", up)):
        sites, rows = [], []
        for rep, gold in zip(gv["replicons"], gv["genes"]):
            genes = [O.OGene(r["start"], r["end"], r["strand"]) for r in gold["rows"]]
            index = [r["index"] for r in gold["rows"]]
            ann = [(r["locus_tag"], r["protein_id"], r["product"]) for r in gold["rows"]]
            ss = O.identify_sites(om, rep["sequence"], genes, use_up, dw, index=index)
            sites.extend((rep, genes, ann, s) for s in ss)
        sites.sort(key=lambda t: t[3].score, reverse=True)          # genome.py:448 over all chromids
        want = gv["scenarios"][scen]["sites"]
        assert len(sites) == len(want)
        for (rep, genes, ann, s), w in zip(sites, want):
            assert (rep["accession"], s.start, s.end, s.strand, float(s.score)) == \
                (w["chromid"], w["start"], w["end"], w["strand"], w["score"])
            assert ann[s.gene][0] == w["gene"]
            rows.extend(O.identified_sites_rows([s], rep["sequence"], genes, up, rep["accession"], ann))
        from cgb_b200.site_search import CSV_HEADER
        assert _csv_text([CSV_HEADER] + rows) == gv["scenarios"][scen]["sites_csv"]


def test_oracle_posteriors_match_reference(gv, lexa):
    om = _oracle_model(gv, lexa)
    for rep, gold, want in zip(gv["replicons"], gv["genes"], gv["regulation_probability"]):
        got = [O.binding_probability(om, r["promoter_region"], gv["prior"], gv["alpha"], softmax="nep50")
               for r in gold["rows"]]
        assert got == want


@pytest.mark.parametrize("name", ["split_0.5_sigma_1.0", "nosplit_sigma_1.0", "split_0.2_sigma_1.5"])
def test_operon_prediction_matches_reference(gv, name):
    """Directons, adaptive distance threshold, posterior-driven splitting, operon ids and the
    two CSV reports (chromid.py:140-243, genome.py:74-152,335-378), from the reference's own
    posteriors."""
    want = gv["operons"][name]
    tables = [_table(g, k) for k, g in enumerate(gv["genes"])]
    accs = [g["accession"] for g in gv["genes"]]
    probs = [np.array(p) for p in gv["regulation_probability"]]
    assert operons.intergenic_distance_threshold(tables, want["sigma"]) == want["distance_threshold"]
    oprs = operons.predict_operons_genome(tables, want["probability_th"], want["sigma"], probs, accs)
    assert len(oprs) == len(want["operons"])
    for o, w in zip(oprs, want["operons"]):
        gt = tables[o.seq_index]
        assert (o.operon_id, o.chromid, o.start, o.end, o.strand) == \
            (w["id"], w["chromid"], w["start"], w["end"], w["strand"])
        assert [gt.locus_tag[g] for g in o.genes] == w["genes"]
        assert gt.locus_tag[o.first_gene] == w["first_gene"]
        assert o.regulation_probability == w["regulation_probability"]
    assert _csv_text(operons.operons_rows(oprs, tables)) == want["operons_csv"]
    results, regulons = operons.infer_regulons(oprs, 0.5)
    assert _csv_text(operons.regulon_rows(results, tables)) == want["regulons_csv"]
    assert [[tables[o.seq_index].locus_tag[g] for g in o.genes] for o in regulons] == want["regulons"]


def test_get_prior_matches_reference(gv):
    """cgb/__init__.py:320-350 over the Genome mirror (host only: IC + operons without splitting)."""
    import types
    from cgb_b200 import Chromid, Genome, get_prior
    chromids = [Chromid(g["accession"], "", rep["sequence"].encode(), _table(g, k))
                for k, (rep, g) in enumerate(zip(gv["replicons"], gv["genes"]))]
    genome = Genome("synthetic_strain", chromids)
    want = gv["inferred_prior"]
    genome._TF_binding_model = types.SimpleNamespace(IC=want["IC"])
    ui = types.SimpleNamespace(prior_regulation_probability=None,
                               operon_prediction_distance_tuning_parameter=want["sigma"])
    assert genome.length == want["genome_length"]
    assert get_prior(genome, ui) == want["prior"]
    assert genome.num_operons == 0                                    # removed again, as the reference does
    ui.prior_regulation_probability = 0.03
    assert get_prior(genome, ui) == 0.03
    with pytest.raises(AttributeError):
        genome.operon_prediction(0.5, 1.0)                            # needs the posteriors first


def test_batched_patser_thresholds_equal_single_calls():
    from cgb_b200 import PSSMModel, SiteCollection, prepare_thresholds
    rng = np.random.default_rng(3)
    models, singles = [], []
    for k in range(24):
        m = int(rng.integers(8, 31))
        cols = rng.dirichlet([0.5] * 4, size=m)
        sites = ["".join("ACGT"[rng.choice(4, p=cols[j])] for j in range(m)) for _ in range(20)]
        singles.append(PSSMModel([SiteCollection(sites)], [1.0]).threshold())
        models.append(PSSMModel([SiteCollection(sites)], [1.0]))
    prepare_thresholds(models, n_threads=4)
    assert all("patser_threshold" in M.__dict__ for M in models)
    assert [M.threshold() for M in models] == singles


def test_genbank_awkward_feature_tables():
    """Multi-line locations, a '/' opening a continuation line inside a quoted value, doubled
    quotes, flag qualifiers, a CONTIG line after the table, lower-case ORIGIN."""
    text = """LOCUS       TEST0001                  60 bp    DNA     linear   BCT 01-JAN-2020
DEFINITION  Awkward test record, with a second
            definition line.
ACCESSION   TEST0001
VERSION     TEST0001.2
FEATURES             Location/Qualifiers
     source          1..60
                     /organism="Test organism"
     gene            complement(join(1..10,
                     20..30))
                     /locus_tag="T_0001"
     gene            11..40
                     /locus_tag="T_0002"
                     /note="a note that wraps and then has a line starting with
                     /slash inside the quotes, and a ""quoted"" word"
                     /pseudo
     CDS             11..40
                     /locus_tag="T_0002"
                     /product="two-line
                     product name"
                     /protein_id="WP_1.1"
                     /translation="MKV
                     LLA"
     gene            41..50
                     /locus_tag="T_0003"
     misc_feature    45..46
CONTIG      join(TEST0001.2:1..60)
ORIGIN
        1 acgtacgtac gtacgtacgt acgtacgtac gtacgtacgt acgtacgtac gtacgtacgn
//
"""
    rec = genbank.read(text)
    assert rec.id == "TEST0001.2" and rec.description == "Awkward test record, with a second definition line"
    assert rec.sequence == b"ACGT" * 14 + b"ACGN" and rec.length == 60
    assert [f.type for f in rec.features] == ["source", "gene", "gene", "CDS", "gene", "misc_feature"]
    g1, g2, cds = rec.features[1], rec.features[2], rec.features[3]
    assert g1.compound and g1.strand == -1
    assert (g2.start, g2.end, g2.strand, g2.compound) == (10, 40, 1, False)
    assert g2.qualifiers["note"] == ['a note that wraps and then has a line starting with /slash inside the '
                                     'quotes, and a "quoted" word']
    assert g2.qualifiers["pseudo"] == [""]
    assert cds.qualifiers["product"] == ["two-line product name"] and cds.qualifiers["translation"] == ["MKVLLA"]
    gt = rec.gene_table()
    assert list(gt.locus_tag) == ["T_0002", "T_0003"] and gt.index.tolist() == [1, 2]   # the join gene is skipped
    assert list(gt.product) == ["two-line product name", ""] and list(gt.protein_id) == ["WP_1.1", ""]


def test_oracle_upstream_background_matches_verbatim_build_PSSM_model(gv, lexa):
    """The background population of Genome.build_PSSM_model (genome.py:215-226): the golden
    estimator was recorded by the reference's own method with random.sample as the identity;
    the oracle's restatement gives the same n and bit-equal mean / std (numpy-2 soft-max mode of
    the recording run) and the pinned-numpy mode within 1e-9."""
    om = O.OracleModel([mo["sites"] for mo in lexa["motifs"]], gv["weights"])
    est = gv["estimator"]
    for mode, tol in (("nep50", 0.0), ("legacy64", 1e-9)):
        bg = []
        for rep, gold in zip(gv["replicons"], gv["genes"]):
            genes = [O.OGene(r["start"], r["end"], r["strand"]) for r in gold["rows"]]
            bg.extend(O.upstream_background_scores(om, rep["sequence"], genes, gv["promoter_dw_distance"],
                                                   [r["index"] for r in gold["rows"]], mode))
        assert len(bg) == est["n_bg"]
        assert abs(np.mean(bg) - est["mu_bg"]) <= tol and abs(np.std(bg) - est["std_bg"]) <= tol
    # the region table the product hands to the kernel covers exactly those m-mers
    m = om.length
    n = 0
    for k, gold in enumerate(gv["genes"]):
        gt = _table(gold, k)
        s, e = python_slice_bounds(*gt.upstream_noncoding_region_location(None, gv["promoter_dw_distance"]),
                                   gold["length"])
        n += int(np.maximum(e - s - m, 0).sum())
    assert n == est["n_bg"]
    # a different population from "all windows of every replicon"
    assert gv["estimator_all_windows"]["n_bg"] != est["n_bg"]
    assert abs(gv["estimator_all_windows"]["mu_bg"] - est["mu_bg"]) > 1e-3


class _OracleScanModel:
    """Stands in for PSSMModel in site_search.identify_sites: the genome-wide hit list comes from
    the oracle, so the vectorised join with the gene regions runs without a GPU."""

    def __init__(self, om, sequences):
        self.om, self.sequences, self.length = om, sequences, om.length

    def threshold(self):
        return self.om.threshold()

    def scan_hits(self, genome, threshold, revseq_order=True):
        seq, pos, strand, score = [], [], [], []
        for k, s in enumerate(self.sequences):
            p, st, v = O.scan_hits_whole(self.om, s, threshold)
            seq.append(np.full(len(p), k)); pos.append(p); strand.append(st); score.append(v)
        return tuple(np.concatenate(x) for x in (seq, pos, strand, score))


class _Lengths:
    def __init__(self, sequences):
        self._len = [len(s) for s in sequences]

    def length(self, k=0):
        return self._len[k]


@pytest.mark.parametrize("scen,use_up", [("scan_to_upstream_gene", False), ("scan_up_distance", True)])
def test_vectorised_site_join_matches_reference(gv, lexa, scen, use_up):
    """site_search.identify_sites (one scan + array join, no per-gene loop) yields the verbatim
    run's sites in the verbatim run's order (genome.py:419-448)."""
    from cgb_b200.site_search import identify_sites
    om = _oracle_model(gv, lexa)
    seqs = [rep["sequence"] for rep in gv["replicons"]]
    tables = [_table(gold, k) for k, gold in enumerate(gv["genes"])]
    sites = identify_sites(_OracleScanModel(om, seqs), _Lengths(seqs), tables, gv["promoter_up_distance"],
                           gv["promoter_dw_distance"], use_up)
    want = gv["scenarios"][scen]["sites"]
    assert len(sites) == len(want)
    for s, w in zip(sites, want):
        assert (gv["replicons"][s.chromid]["accession"], s.start, s.end, s.strand, float(s.score),
                tables[s.chromid].locus_tag[s.gene]) == \
            (w["chromid"], w["start"], w["end"], w["strand"], w["score"], w["gene"])


def test_site_join_raises_like_biopython_on_too_short_region():
    from cgb_b200.site_search import identify_sites

    class M:
        length = 22
        def threshold(self):
            return 5.0
        def scan_hits(self, genome, threshold, revseq_order=True):
            z = np.zeros(0, dtype=np.int64)
            return z, z, z.astype(np.int32), z.astype(np.float32)

    # forward gene at 5 with down = 10: region [0, 15), 15 < m - 1 = 21 -> numpy.empty(negative)
    gt = GeneTable([5], [400], [1], 1000)
    with pytest.raises(ValueError):
        identify_sites(M(), _Lengths(["A" * 1000]), [gt], 300, 10, False)
    # a region of exactly m - 1 bases scores no window and raises nothing
    gt = GeneTable([11], [400], [1], 1000)
    assert identify_sites(M(), _Lengths(["A" * 1000]), [gt], 300, 10, False) == []

"""Randomised differential tests against the reference run VERBATIM (build container only: they
need /root/reference and are skipped elsewhere, e.g. on the GPU box).  Where the committed golden
fixtures pin one genome, these throw many random ones at the same code: random feature tables
(compound genes, missing product features, overlapping genes, a trailing gene) through
Chromid.genes / Gene.* / Chromid.operon_prediction on one side and cgb_b200.genbank / GeneTable /
operons on the other; random sequences and motifs through PSSMModel.score_seq /
binding_probability against the oracle."""
import types

import numpy as np
import pytest

from oracle import ref_loader
from oracle import np_oracle as O

pytestmark = pytest.mark.skipif(not ref_loader.available(), reason="reference tree not present")


def _random_replicon(rng, accession, length):
    feats = [{"type": "source", "start": 0, "end": length, "strand": 1, "compound": False, "qualifiers": {}}]
    pos, strand, k = int(rng.integers(0, 300)), int(rng.choice([1, -1])), 0
    while True:
        glen = int(rng.integers(60, 900))
        if pos + glen > length:
            break
        k += 1
        tag = "%s_%04d" % (accession[:3], k)
        kind = rng.choice(["cds", "cds_noid", "trna", "none", "misc", "compound"], p=[0.55, 0.1, 0.1, 0.1, 0.05, 0.1])
        feats.append({"type": "gene", "start": pos, "end": pos + glen, "strand": strand,
                      "compound": bool(kind == "compound"), "qualifiers": {"locus_tag": [tag]}})
        if kind in ("cds", "cds_noid", "compound"):
            q = {"locus_tag": [tag], "product": ["p%d" % k]}
            if kind != "cds_noid":
                q["protein_id"] = ["WP_%d.1" % k]
            feats.append({"type": "CDS", "start": pos, "end": pos + glen, "strand": strand,
                          "compound": bool(kind == "compound"), "qualifiers": q})
        elif kind == "trna":
            feats.append({"type": "tRNA", "start": pos, "end": pos + glen, "strand": strand, "compound": False,
                          "qualifiers": {"locus_tag": [tag], "product": ["tRNA-Gly"]}})
        elif kind == "misc":
            feats.append({"type": "misc_feature", "start": pos, "end": pos + 10, "strand": strand, "compound": False,
                          "qualifiers": {}})
        if rng.random() < 0.5:
            gap = int(rng.integers(-15, 40))
        else:
            gap = int(rng.integers(60, 500))
            if rng.random() < 0.5:
                strand = -strand
        pos = max(0, pos + glen + gap)
    seq = "".join(rng.choice(list("ACGT"), size=length))
    return {"accession": accession, "description": "random", "sequence": seq, "features": feats}


def _record(rep):
    feats = []
    for f in rep["features"]:
        loc = (ref_loader.CompoundLocation([], f["strand"]) if f["compound"]
               else ref_loader.FeatureLocation(f["start"], f["end"], f["strand"]))
        feats.append(ref_loader.FakeFeature(f["type"], loc, {k: list(v) for k, v in f["qualifiers"].items()}))
    return ref_loader.FakeRecord(rep["accession"], rep["description"], rep["sequence"], feats)


@pytest.mark.parametrize("seed", range(12))
def test_gene_tables_coordinates_and_operons(seed):
    from cgb_b200 import genbank, operons, python_slice_bounds
    from genbank_writer import genbank_text
    ns = ref_loader.load_genome_stack()
    rng = np.random.default_rng(1000 + seed)
    reps = [_random_replicon(rng, "R%02dA.1" % seed, int(rng.integers(8000, 30000))),
            _random_replicon(rng, "R%02dB.1" % seed, int(rng.integers(3000, 9000)))]
    for rep in reps:
        ref_loader.register_record(_record(rep))
    genome = ns.Genome("random_%d" % seed, [rep["accession"] for rep in reps])
    tables = []
    up, dw = int(rng.integers(50, 400)), int(rng.integers(0, 80))
    for k, (rep, chromid) in enumerate(zip(reps, genome.chromids)):
        gt = genbank.read(genbank_text(rep, fuzzy_first=False)).gene_table(k)
        tables.append(gt)
        genes = chromid.genes
        assert gt.start.tolist() == [g.start for g in genes] and gt.end.tolist() == [g.end for g in genes]
        assert gt.strand.tolist() == [g.strand for g in genes] and gt.index.tolist() == [g._index for g in genes]
        assert list(gt.locus_tag) == [g.locus_tag for g in genes]
        assert list(gt.product_type) == [g.product_type for g in genes] and list(gt.product) == [g.product for g in genes]
        try:
            want_none = [list(g.upstream_noncoding_region_location(None, dw)) for g in genes]
        except IndexError:
            with pytest.raises(IndexError):
                gt.upstream_noncoding_region_location(None, dw)
        else:
            s, e = gt.upstream_noncoding_region_location(None, dw)
            assert [[int(a), int(b)] for a, b in zip(s, e)] == want_none
        s, e = gt.upstream_noncoding_region_location(up, dw)
        assert [[int(a), int(b)] for a, b in zip(s, e)] == [list(g.upstream_noncoding_region_location(up, dw))
                                                            for g in genes]
        s, e = gt.promoter_region_location(up, dw)
        s, e = python_slice_bounds(s, e, len(rep["sequence"]))
        assert [rep["sequence"][int(a):int(b)] for a, b in zip(s, e)] == [g.promoter_region(up, dw) for g in genes]
    # operons: random posteriors on the genes, three settings of (threshold, sigma)
    probs = []
    for chromid in genome.chromids:
        p = rng.random(len(chromid.genes)) ** 3
        for g, x in zip(chromid.genes, p):
            g._regulation_probability = float(x)
        probs.append(p)
    if not any(len(d) > 1 for d in genome.directons):
        return                                        # the reference divides by zero here
    for th, sigma in ((0.5, 1.0), (1.0, 1.0), (0.1, 2.0)):
        genome.operon_prediction(th, sigma)
        got = operons.predict_operons_genome(tables, th, sigma, probs, [r["accession"] for r in reps])
        want = genome.operons
        assert operons.intergenic_distance_threshold(tables, sigma) == genome.intergenic_distance_threshold(sigma)
        assert [(o.operon_id, o.chromid, o.start, o.end, o.strand) for o in got] == \
            [(o.operon_id, o.chromid.accession_number, o.start, o.end, o.strand) for o in want]
        assert [[tables[o.seq_index].locus_tag[g] for g in o.genes] for o in got] == \
            [[g.locus_tag for g in o.genes] for o in want]
        assert [o.regulation_probability for o in got] == [o.regulation_probability for o in want]


@pytest.mark.parametrize("seed", range(6))
def test_oracle_scores_and_posteriors_against_verbatim_model(seed, lexa_site_lists):
    """Random weightings of the LexA collections, random sequences with ambiguity codes and
    lower case: the oracle (nep50 mode = this container's numpy) equals the verbatim PSSMModel."""
    ns = ref_loader.load()
    rng = np.random.default_rng(2000 + seed)
    colls = ref_loader.lexa_collections(ns)
    w = rng.dirichlet([1.0] * len(colls)).tolist()
    M = ns.PSSMModel(colls, w)
    om = O.OracleModel(lexa_site_lists, w)
    assert M.IC == om.IC and M.threshold() == om.threshold()
    seq = "".join(rng.choice(list("ACGT"), size=1500))
    seq = seq[:300] + "N" + seq[301:700] + seq[700:760].lower() + seq[760:1200] + "RYK" + seq[1203:]
    got = M.score_seq(seq)
    exp = O.score_seq(om, seq, True, "nep50")[0]
    assert [float(x) for x in got] == [float(x) for x in exp]
    f = M.score_seq(seq, both=False)
    assert np.array_equal(np.asarray(f, dtype=np.float64), np.asarray(O.score_seq(om, seq, both=False)[0],
                                                                      dtype=np.float64), equal_nan=True)
    bg = M.score_seq(seq)
    M.build_bayesian_estimator(bg)
    O.build_bayesian_estimator(om, exp, softmax="nep50")
    for lo in (0, 400, 1100):
        assert M.binding_probability(seq[lo:lo + 350], 0.03, 1 / 350.0) == \
            O.binding_probability(om, seq[lo:lo + 350], 0.03, 1 / 350.0, softmax="nep50")

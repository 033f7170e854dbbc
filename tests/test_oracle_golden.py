"""Pin the oracle: oracle/np_oracle.py and oracle/pwm_oracle.c against the golden
vectors recorded from the reference's own pssm_model.py / binding_model.py run
verbatim (tests/golden/make_golden.py), plus the hand-derivable KAT of SURVEY §8c."""
import math

import numpy as np
import pytest

from oracle import c_oracle, np_oracle as O


def _model(lexa_site_lists, weights):
    return O.OracleModel(lexa_site_lists, weights)


@pytest.mark.parametrize("wname", ["equal", "site_count"])
def test_matrices_ic_threshold(golden, lexa_site_lists, lexa_weights, wname):
    rec = golden["models"][wname]
    m = _model(lexa_site_lists, lexa_weights[wname])
    for l in "ACGT":
        assert list(m.pwm[l]) == rec["pwm"][l]
        assert m.pssm[l] == rec["pssm"][l]
        assert m.rev_comp_pssm[l] == rec["rev_comp_pssm"][l]
    assert m.IC == rec["IC"]
    assert m.length == rec["length"] == 22
    assert m.threshold() == rec["patser_threshold"]
    with pytest.raises(ValueError):
        m.threshold("other")
    assert rec["threshold_other"] == "ValueError"


def test_patser_literal_matches_vectorised(lexa_site_lists):
    m = O.OracleModel([lexa_site_lists[3]], [1.0])
    assert O.patser_threshold(m.pssm, precision=50, literal=True) == O.patser_threshold(m.pssm, precision=50)


def test_survey_values(golden):
    # SURVEY.md §8c "survey-derived vectors", 6 decimals
    exp = {"LexA_Mtu": (18.583570, 15.045293), "LexA_Cgl": (12.071128, 8.689919),
           "LexA_Lmo": (14.442433, 10.947171), "LexA_Sau": (11.784797, 8.359238),
           "LexA_Bsu": (16.012572, 12.549741), "LexA_Cdi": (10.175328, 6.865170)}
    for c in golden["collections"]:
        ic, thr = exp[c["name"]]
        assert abs(c["IC"] - ic) < 5e-7 and abs(c["patser_threshold"] - thr) < 5e-7
    assert abs(golden["models"]["site_count"]["IC"] - 12.388789) < 5e-7
    assert abs(golden["models"]["site_count"]["patser_threshold"] - 8.988094) < 5e-7


def test_per_collection(golden, lexa_site_lists):
    for c, sites in zip(golden["collections"], lexa_site_lists):
        m = O.OracleModel([sites], [1.0])
        for l in "ACGT":
            assert list(m.collection_pwms[0][l]) == c["pwm"][l]
        assert m.IC == c["IC"]
        assert m.threshold() == c["patser_threshold"]
        s, r = O.score_seq(m, c["first_site"], both=False)
        assert float(s[0]) == c["first_site_fwd"] and float(r[0]) == c["first_site_rc"]
        assert O.score_seq(m, c["first_site"], True, "nep50")[0][0] == c["first_site_both"]


def test_kat_acgt(golden):
    k = golden["kat_acgt"]
    m = O.OracleModel([["ACGT"]], [1.0])
    assert m.pwm["A"] == (0.4, 0.2, 0.2, 0.2)
    assert abs(m.pssm["A"][0] - math.log2(1.6)) < 1e-15 and abs(m.pssm["C"][0] - math.log2(0.8)) < 1e-15
    f = O.score_seq(m, "ACGT", both=False)[0][0]
    assert float(f) == k["fwd"] == float(np.float32(4 * math.log2(1.6)))
    assert O.score_seq(m, "ACGT", True, "nep50")[0][0] == k["both"]
    assert abs(O.score_seq(m, "ACGT")[0][0] - 3.7122876) < 1e-6


@pytest.mark.parametrize("wname", ["equal", "site_count"])
def test_score_seq_bit_exact(golden, lexa_site_lists, lexa_weights, wname):
    rec = golden["models"][wname]
    m = _model(lexa_site_lists, lexa_weights[wname])
    for cname in ("random2k", "with_N", "lowercase"):
        case = rec["score_seq"][cname]
        both, _ = O.score_seq(m, case["seq"], True, "nep50")
        assert both == case["both"], cname
        fwd, _ = O.score_seq(m, case["seq"], False)
        assert [float(x) for x in fwd] == case["fwd_f32"], cname
        # the semantics the CUDA path implements (pinned numpy 1.x) differ by < 2e-7
        legacy, _ = O.score_seq(m, case["seq"], True)
        assert np.max(np.abs(np.array(legacy) - np.array(case["both"]))) < 5e-7
        # C restatement == numpy restatement
        f, r, b = c_oracle.score_seq(m, case["seq"].encode())
        assert np.array_equal(f, np.asarray(fwd, dtype=np.float32))
        assert np.max(np.abs(b - np.array(legacy))) < 1e-12
        vec, _, _ = O.score_seq_both_vec(m, case["seq"])
        assert np.max(np.abs(vec - np.array(legacy))) < 1e-12
    for s, v in rec["score_seq"]["single"].items():
        assert O.score_seq(m, s, True, "nep50")[0][0] == v["both"]
        assert float(O.score_seq(m, s, False)[0][0]) == v["fwd"]
    assert [x[0] for x in O.score_self(m, "nep50")] == rec["score_self"]


@pytest.mark.parametrize("wname", ["equal", "site_count"])
def test_estimator_and_posteriors(golden, lexa_site_lists, lexa_weights, wname):
    rec = golden["models"][wname]
    m = _model(lexa_site_lists, lexa_weights[wname])
    seq = rec["score_seq"]["random2k"]["seq"]
    bg = [[x] for x in O.score_seq(m, seq, True, "nep50")[0]]
    O.build_bayesian_estimator(m, bg, "nep50")
    est = rec["estimator"]
    assert (m._mu_m, m._std_m, m._mu_bg, m._std_bg) == (est["mu_m"], est["std_m"], est["mu_bg"], est["std_bg"])
    for p in rec["posteriors"]:
        got = O.binding_probability(m, p["seq"], p["p_motif"], p["alpha"], "nep50")
        assert abs(got - p["posterior"]) <= 1e-12 * max(1.0, abs(p["posterior"]))
        legacy = O.binding_probability(m, p["seq"], p["p_motif"], p["alpha"])
        assert abs(legacy - p["posterior"]) < 1e-6
        c = c_oracle.binding_probability(m, p["seq"].encode(), m._mu_m, m._std_m, m._mu_bg, m._std_bg,
                                         p["p_motif"], p["alpha"])
        assert abs(c - legacy) < 1e-9


def test_identify_sites_matches_c(lexa_site_lists, lexa_weights):
    m = _model(lexa_site_lists, lexa_weights["site_count"])
    seq = O.synthetic_genome(60000, 3, n_frac=0.002)
    thr = m.threshold() - 3.0
    pos, strand, score = O.scan_hits_whole(m, seq, thr)
    cpos, cstrand, cscore = c_oracle.scan_hits(m, seq, thr)
    assert len(pos) > 20
    assert np.array_equal(pos, cpos) and np.array_equal(strand, cstrand) and np.array_equal(score, cscore)
    genes = O.synthetic_genes(len(seq), 30)
    sites = O.identify_sites(m, seq.decode(), genes, None, 50, thr)
    assert all(sites[i].score >= sites[i + 1].score for i in range(len(sites) - 1))
    whole = set(zip(pos.tolist(), strand.tolist()))
    assert all((s.start, s.strand) in whole for s in sites)


def test_coordinate_rules():
    genes = [O.OGene(100, 400, 1), O.OGene(420, 900, 1), O.OGene(1000, 1500, -1), O.OGene(1600, 1990, -1)]
    L = 2000
    assert O.upstream_noncoding_region_location(genes, 0, L) == (0, 150)
    assert O.upstream_noncoding_region_location(genes, 1, L) == (350, 470)
    assert O.upstream_noncoding_region_location(genes, 2, L) == (1450, 1650)
    assert O.upstream_noncoding_region_location(genes, 3, L) == (1940, 2000)
    assert O.upstream_noncoding_region_location(genes, 1, L, up=300) == (120, 470)
    assert O.upstream_noncoding_region_location(genes, 1, L, up=0) == (350, 470)   # `if up:`
    assert O.promoter_region_location(genes[0], L) == (0, 150)
    assert O.promoter_region_location(genes[3], L) == (1940, 2000)
    assert O.promoter_region_location(genes[2], L, 300, 50) == (1450, 1800)
